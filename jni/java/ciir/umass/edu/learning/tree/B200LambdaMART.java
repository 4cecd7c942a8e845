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

        if (validationSamples != null) {
            modelScoresOnValidation = new double[validationSamples.size()][];
            for (int i = 0; i < validationSamples.size(); i++) {
                modelScoresOnValidation[i] = new double[validationSamples.get(i).size()];
            }
        }

        release();
        handle = NativeBridge.create(device);
        NativeBridge.loadDense(handle, x, n, nf, features, labels, qoff);
        NativeBridge.init(handle, nTreeLeaves, minLeafSupport, learningRate, nThreshold, kind(), metricCode(scorer),
                scorer.getK(), FeatureHistogram.samplingRate, seed);
    }

    /**
     * init() for a training set that is still a file: FeatureManager.readInput (FeatureManager.java:187-245) and the
     * flattening above happen in the library's multithreaded reader, and no DataPoint object is created.  `samples` stays
     * empty, so the final scorer.score(rank(samples)) of learn() is replaced by the library's own score on the training
     * data (scoreOnTrainingData keeps the last NDCG@k-T); validation sets still go through setValidationSet.
     */
    public void initFromFile(final String trainingFile, final boolean mustHaveRelDoc) {
        if (validationSamples != null) {
            modelScoresOnValidation = new double[validationSamples.size()][];
            for (int i = 0; i < validationSamples.size(); i++) {
                modelScoresOnValidation[i] = new double[validationSamples.get(i).size()];
            }
        }
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
        NativeBridge.init(handle, nTreeLeaves, minLeafSupport, learningRate, nThreshold, kind(), metricCode(scorer),
                scorer.getK(), FeatureHistogram.samplingRate, seed);
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
                printLog(new int[] { 9 }, new String[] { Double.toString(SimpleMath.round(scoreOnTrainingData, 4)) });

                if (validationSamples != null) {
                    for (int i = 0; i < modelScoresOnValidation.length; i++) {
                        final RankList rl = validationSamples.get(i);
                        for (int j = 0; j < modelScoresOnValidation[i].length; j++) {
                            modelScoresOnValidation[i][j] += learningRate * rt.eval(rl.get(j));
                        }
                    }
                    final double v = computeModelScoreOnValidation();
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
        } finally {
            release();
        }

        while (ensemble.treeCount() > bestModelOnValidation + 1) {
            ensemble.remove(ensemble.treeCount() - 1);
        }

        if (!samples.isEmpty()) { // initFromFile leaves no Java-side copy of the training set
            scoreOnTrainingData = scorer.score(rank(samples));
        }
        if (validationSamples != null) {
            bestScoreOnValidationData = scorer.score(rank(validationSamples));
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
