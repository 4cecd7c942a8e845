/*
 * B200LambdaMART — LambdaMART whose init() and boosting loop run in libranklib_b200.so.
 *
 * A subclass, so no file of the reference has to be edited: every protected field it touches
 * (samples, features, scorer, validationSamples, martSamples, modelScores, modelScoresOnValidation, ensemble,
 * bestModelOnValidation, scoreOnTrainingData, bestScoreOnValidationData) belongs to Ranker / LambdaMART
 * (R/learning/Ranker.java:39-46, R/learning/tree/LambdaMART.java:45-58).  Create it through the factory's
 * class-name route (R/learning/RankerFactory.java:72-94):
 *
 *     Ranker r = new RankerFactory().createRanker("ciir.umass.edu.learning.tree.B200LambdaMART", train, features, scorer);
 *     r.setValidationSet(vali); r.init(); r.learn(); r.save("model.txt");
 *
 * name() stays "LambdaMART": the model text (model(), Ensemble.toString()) is the reference's and loads in stock
 * RankLib.  NOT COMPILED IN THIS IMAGE (no JDK); the Python mirror ranklib_b200/host/rankers.py:LambdaMART is the
 * executable statement of the same control flow and is what tests/test_host_mirror.py runs.
 */
package ciir.umass.edu.learning.tree;

import java.util.List;

import ciir.umass.edu.learning.DataPoint;
import ciir.umass.edu.learning.RankList;
import ciir.umass.edu.learning.Ranker;
import ciir.umass.edu.metric.APScorer;
import ciir.umass.edu.metric.BestAtKScorer;
import ciir.umass.edu.metric.DCGScorer;
import ciir.umass.edu.metric.ERRScorer;
import ciir.umass.edu.metric.MetricScorer;
import ciir.umass.edu.metric.NDCGScorer;
import ciir.umass.edu.metric.PrecisionScorer;
import ciir.umass.edu.metric.ReciprocalRankScorer;
import ciir.umass.edu.utilities.RankLibError;
import ciir.umass.edu.utilities.SimpleMath;

public class B200LambdaMART extends LambdaMART {
    /** CUDA device of this ranker's native context. */
    public static int device = 0;
    /** Seed of the feature-sampling stream (the reference's is unseeded, FeatureHistogram.java:282). */
    public static long seed = 0L;

    protected long handle = 0L;

    public B200LambdaMART() {
    }

    public B200LambdaMART(final List<RankList> samples, final int[] features, final MetricScorer scorer) {
        super(samples, features, scorer);
    }

    /** RLB_KIND_*: which pseudo responses / leaf outputs the library computes. */
    protected int kind() {
        return NativeBridge.KIND_LAMBDAMART;
    }

    /** RLB_METRIC_* of the training scorer (NDCGScorer extends DCGScorer: test the subclass first). */
    protected static int metricCode(final MetricScorer s) {
        if (s instanceof NDCGScorer) {
            return NativeBridge.METRIC_NDCG;
        } else if (s instanceof DCGScorer) {
            return NativeBridge.METRIC_DCG;
        } else if (s instanceof ERRScorer) {
            return NativeBridge.METRIC_ERR;
        } else if (s instanceof APScorer) {
            return NativeBridge.METRIC_MAP;
        } else if (s instanceof PrecisionScorer) {
            return NativeBridge.METRIC_PRECISION;
        } else if (s instanceof ReciprocalRankScorer) {
            return NativeBridge.METRIC_RR;
        } else if (s instanceof BestAtKScorer) {
            return NativeBridge.METRIC_BEST;
        }
        throw RankLibError.create("B200LambdaMART: no device implementation of the metric " + s.name());
    }

    /**
     * Replaces LambdaMART.init (LambdaMART.java:68-166): the flattening of the rank lists stays here, the per-feature
     * sorts, the candidate thresholds and FeatureHistogram.construct happen on the device.
     */
    @Override
    public void init() {
        int n = 0;
        for (final RankList rl : samples) {
            n += rl.size();
        }
        final int nf = features.length;
        final float[] x = new float[Math.multiplyExact(n, nf)];
        final float[] labels = new float[n];
        final int[] qoff = new int[samples.size() + 1];
        martSamples = new DataPoint[n];
        modelScores = new double[n];
        impacts = new double[nf]; // never updated by the reference either (SURVEY.md Q1); RFRanker.learn reads it
        int at = 0;
        for (int q = 0; q < samples.size(); q++) {
            final RankList rl = samples.get(q);
            qoff[q] = at;
            for (int j = 0; j < rl.size(); j++, at++) {
                final DataPoint dp = rl.get(j);
                martSamples[at] = dp;
                labels[at] = dp.getLabel();
                final int row = at * nf;
                for (int c = 0; c < nf; c++) {
                    x[row + c] = dp.getFeatureValue(features[c]); // unknown -> 0, DenseDataPoint.java:21-32
                }
            }
        }
        qoff[samples.size()] = at;

        release();
        handle = NativeBridge.create(device);
        NativeBridge.loadDense(handle, x, n, nf, features, labels, qoff);
        uploadValidation();
        NativeBridge.init(handle, nTreeLeaves, minLeafSupport, learningRate, nThreshold, kind(), metricCode(scorer),
                scorer.getK(), FeatureHistogram.samplingRate, seed);
    }

    /**
     * init() for a bag of a Random Forest (B200RFRanker): the training set of this ranker is gathered on the device from
     * the context `source`, which holds the forest's whole training set — Sampler.doSampling without host work.
     */
    void initFromBag(final long source, final int[] picks, final int nDocs) {
        release();
        handle = NativeBridge.create(device);
        NativeBridge.loadBag(handle, source, picks);
        modelScores = new double[nDocs];
        impacts = new double[features.length];
        NativeBridge.init(handle, nTreeLeaves, minLeafSupport, learningRate, nThreshold, kind(), metricCode(scorer),
                scorer.getK(), FeatureHistogram.samplingRate, seed);
    }

    /**
     * modelScoresOnValidation of LambdaMART.init (LambdaMART.java:152-158): the validation lists go to the device once, in
     * the training set's feature columns; the per-tree update and metric (LambdaMART.java:228-237) then run there.
     */
    protected void uploadValidation() {
        if (validationSamples == null) {
            return;
        }
        int n = 0;
        for (final RankList rl : validationSamples) {
            n += rl.size();
        }
        final int nf = features.length;
        final float[] x = new float[Math.multiplyExact(n, nf)];
        final float[] labels = new float[n];
        final int[] qoff = new int[validationSamples.size() + 1];
        int at = 0;
        for (int q = 0; q < validationSamples.size(); q++) {
            final RankList rl = validationSamples.get(q);
            qoff[q] = at;
            for (int j = 0; j < rl.size(); j++, at++) {
                final DataPoint dp = rl.get(j);
                labels[at] = dp.getLabel();
                for (int c = 0; c < nf; c++) {
                    x[at * nf + c] = dp.getFeatureValue(features[c]);
                }
            }
        }
        qoff[validationSamples.size()] = at;
        NativeBridge.loadValidation(handle, x, n, nf, labels, qoff);
    }

    /**
     * init() for a training set that is still a file: FeatureManager.readInput (FeatureManager.java:187-245) and the
     * flattening above happen in the library's multithreaded reader, and no DataPoint object is created.  `samples` stays
     * empty, so the final scorer.score(rank(samples)) of learn() is replaced by the library's own score on the training
     * data (scoreOnTrainingData keeps the last NDCG@k-T); validation sets still go through setValidationSet.
     */
    public void initFromFile(final String trainingFile, final boolean mustHaveRelDoc) {
        release();
        handle = NativeBridge.create(device);
        final int[] dims = new int[3];
        NativeBridge.loadLetorFile(handle, trainingFile, mustHaveRelDoc, features, dims);
        if (features == null) {
            features = new int[dims[2]];
            for (int f = 0; f < features.length; f++) {
                features[f] = f + 1;
            }
        }
        modelScores = new double[dims[0]];
        impacts = new double[features.length];
        uploadValidation();
        NativeBridge.init(handle, nTreeLeaves, minLeafSupport, learningRate, nThreshold, kind(), metricCode(scorer),
                scorer.getK(), FeatureHistogram.samplingRate, seed);
    }

    /** Flat node arrays of every tree of the ensemble, as NativeBridge.boostIter returned them (Split keeps its feature id and
     * threshold private, so the arrays the device path needs are kept beside the Split objects). */
    protected final java.util.List<int[]> treeInts = new java.util.ArrayList<>();
    protected final java.util.List<float[]> treeFloats = new java.util.ArrayList<>();

    /** scorer.score(rank(samples)) from the set resident on the device: which = 0 training, 1 validation; the first
     * ensemble.treeCount() trees (the roll-back of LambdaMART.java:254-256 has already dropped the others). */
    protected double scoreResident(final int which) {
        final int nt = ensemble.treeCount();
        final int[] off = new int[nt + 1];
        final float[] w = new float[nt];
        int total = 0;
        for (int t = 0; t < nt; t++) {
            off[t] = total;
            w[t] = ensemble.getWeight(t);
            total += treeInts.get(t).length / 7;
        }
        off[nt] = total;
        final int[] ni = new int[7 * total];
        final float[] nf = new float[2 * total];
        for (int t = 0; t < nt; t++) {
            System.arraycopy(treeInts.get(t), 0, ni, 7 * off[t], treeInts.get(t).length);
            System.arraycopy(treeFloats.get(t), 0, nf, 2 * off[t], treeFloats.get(t).length);
        }
        return NativeBridge.scoreResident(handle, which, ni, nf, off, w);
    }

    /** Split objects (Split.java:44-83) from the flat node arrays of NativeBridge.boostIter. */
    protected static Split treeFromFlat(final int[] ni, final float[] nf, final double[] nd, final int node) {
        final int left = ni[7 * node + 3];
        if (left < 0) {
            final Split leaf = new Split();
            leaf.setOutput(nf[2 * node + 1]);
            return leaf;
        }
        final Split s = new Split(ni[7 * node], nf[2 * node], nd[node]);
        s.setLeft(treeFromFlat(ni, nf, nd, left));
        s.setRight(treeFromFlat(ni, nf, nd, ni[7 * node + 4]));
        return s;
    }

    /**
     * The loop of LambdaMART.learn (LambdaMART.java:169-272) with its body (:183-216: pseudo responses, histogram
     * update, tree fit, leaf outputs, score update, training metric) replaced by ONE native call per tree.  Logging,
     * validation scoring, best-model tracking, early stop and roll-back keep the reference's rules.
     */
    @Override
    public void learn() {
        ensemble = new Ensemble();
        treeInts.clear();
        treeFloats.clear();
        final int cap = NativeBridge.nodeCapacity(nTreeLeaves);
        final int[] ni = new int[7 * cap];
        final float[] nf = new float[2 * cap];
        final double[] nd = new double[cap];
        final int[] nn = new int[1];

        if (validationSamples != null) {
            printLogLn(new int[] { 7, 9, 9 }, new String[] { "#iter", scorer.name() + "-T", scorer.name() + "-V" });
        } else {
            printLogLn(new int[] { 7, 9 }, new String[] { "#iter", scorer.name() + "-T" });
        }

        try {
            for (int m = 0; m < nTrees; m++) {
                printLog(new int[] { 7 }, new String[] { Integer.toString(m + 1) });

                scoreOnTrainingData = NativeBridge.boostIter(handle, ni, nf, nd, nn);
                final RegressionTree rt = new RegressionTree(treeFromFlat(ni, nf, nd, 0));
                ensemble.add(rt, learningRate);
                treeInts.add(java.util.Arrays.copyOf(ni, 7 * nn[0]));
                treeFloats.add(java.util.Arrays.copyOf(nf, 2 * nn[0]));
                printLog(new int[] { 9 }, new String[] { Double.toString(SimpleMath.round(scoreOnTrainingData, 4)) });

                if (validationSamples != null) {
                    // LambdaMART.java:228-237 ran on the device inside boostIter (resident validation lists)
                    final double v = NativeBridge.validMetric(handle);
                    printLog(new int[] { 9 }, new String[] { Double.toString(SimpleMath.round(v, 4)) });
                    if (v > bestScoreOnValidationData) {
                        bestScoreOnValidationData = v;
                        bestModelOnValidation = ensemble.treeCount() - 1;
                    }
                }
                flushLog();

                if (m - bestModelOnValidation > nRoundToStopEarly) {
                    break;
                }
            }
            NativeBridge.readScores(handle, modelScores);

            while (ensemble.treeCount() > bestModelOnValidation + 1) {
                ensemble.remove(ensemble.treeCount() - 1);
            }
            // scorer.score(rank(samples)) (LambdaMART.java:259,263) from the matrices already on the device
            scoreOnTrainingData = scoreResident(0);
            if (validationSamples != null) {
                bestScoreOnValidationData = scoreResident(1);
            }
        } finally {
            release();
        }
    }

    /** Frees the device memory of this ranker's context. */
    protected void release() {
        if (handle != 0L) {
            final long h = handle;
            handle = 0L;
            NativeBridge.destroy(h);
        }
    }

    @Override
    public Ranker createNew() {
        return new B200LambdaMART();
    }
}
