/*
 * B200RFRanker — Random Forests (-ranker 8) whose bags are trained by the native path.
 *
 * RFRanker.learn (R/learning/tree/RFRanker.java:72-114) builds every bag with
 * rf.createRanker(rType, bag, features, scorer), i.e. a stock MART / LambdaMART: with only B200LambdaMART / B200MART
 * registered, -ranker 8 would never reach the GPU.  This subclass keeps RFRanker's parameters, model text, eval() and
 * rank(), and replaces learn(): the training set is uploaded ONCE; every bag is Sampler.doSampling's pick of lists
 * (with replacement, R/learning/Sampler.java:21-38) gathered on the device (NativeBridge.loadBag), initialised
 * (the bag's own thresholds, LambdaMART.java:94-150) and fitted there by a B200MART / B200LambdaMART.
 *
 * The reference's Sampler and the per-split feature sampling draw from unseeded `new Random()` (Sampler.java:22,
 * FeatureHistogram.java:282); here both come from java.util.Random streams seeded from `seed`, so that a run can be
 * reproduced and compared (bag i: picks from Random(seed) in bag order, feature sampling seeded seed + 1 + i — the same
 * convention as ranklib_b200/host/rankers.py:RFRanker and the oracle).
 *
 * NOT COMPILED IN THIS IMAGE (no JDK, SURVEY.md F1); the natives it calls are executed under the mock JNIEnv
 * (tests/test_zz_jni_shim.py::test_shim_device_side_bag_equals_ctypes_binding).
 */
package ciir.umass.edu.learning.tree;

import java.util.List;
import java.util.Random;

import ciir.umass.edu.learning.DataPoint;
import ciir.umass.edu.learning.RankerType;
import ciir.umass.edu.learning.RankList;
import ciir.umass.edu.learning.Ranker;
import ciir.umass.edu.metric.MetricScorer;
import ciir.umass.edu.utilities.SimpleMath;

public class B200RFRanker extends RFRanker {
    /** Seed of the bagging and feature-sampling streams. */
    public static long seed = 0L;

    public B200RFRanker() {
    }

    public B200RFRanker(final List<RankList> samples, final int[] features, final MetricScorer scorer) {
        super(samples, features, scorer);
    }

    @Override
    public void learn() {
        printLogLn(new int[] { 9, 9, 11 }, new String[] { "bag", scorer.name() + "-B", scorer.name() + "-OOB" });
        // the whole training set on the device, once (what LambdaMART.init flattens, LambdaMART.java:71-91)
        int n = 0;
        for (final RankList rl : samples) {
            n += rl.size();
        }
        final int nf = features.length;
        final float[] x = new float[Math.multiplyExact(n, nf)];
        final float[] labels = new float[n];
        final int[] qoff = new int[samples.size() + 1];
        int at = 0;
        for (int q = 0; q < samples.size(); q++) {
            final RankList rl = samples.get(q);
            qoff[q] = at;
            for (int j = 0; j < rl.size(); j++, at++) {
                final DataPoint dp = rl.get(j);
                labels[at] = dp.getLabel();
                for (int c = 0; c < nf; c++) {
                    x[at * nf + c] = dp.getFeatureValue(features[c]);
                }
            }
        }
        qoff[samples.size()] = at;
        final long base = NativeBridge.create(B200LambdaMART.device);
        try {
            NativeBridge.loadDense(base, x, n, nf, features, labels, qoff);
            final Random bagging = new Random(seed);
            final int size = (int) (subSamplingRate * samples.size());   // Sampler.doSampling: samplingRate * samplingPool.size()
            double[] impacts = null;
            for (int i = 0; i < nBag; i++) {
                final int[] picks = new int[size];
                int docs = 0;
                for (int k = 0; k < size; k++) {                          // with replacement (Sampler.java:27-31)
                    picks[k] = bagging.nextInt(samples.size());
                    docs += samples.get(picks[k]).size();
                }
                final B200LambdaMART r = (rType == RankerType.MART) ? new B200MART() : new B200LambdaMART();
                r.setFeatures(features);
                r.setMetricScorer(scorer);
                B200LambdaMART.seed = seed + 1 + i;
                r.initFromBag(base, picks, docs);
                r.learn();
                if (impacts == null) {
                    impacts = r.impacts;
                } else {
                    for (int ftr = 0; ftr < impacts.length; ftr++) {
                        impacts[ftr] += r.impacts[ftr];
                    }
                }
                printLogLn(new int[] { 9, 9 }, new String[] { "b[" + (i + 1) + "]", SimpleMath.round(r.getScoreOnTrainingData(), 4) + "" });
                ensembles[i] = r.getEnsemble();
            }
        } finally {
            NativeBridge.destroy(base);
        }
        // RFRanker.java:97-103
        scoreOnTrainingData = scorer.score(rank(samples));
        if (validationSamples != null) {
            bestScoreOnValidationData = scorer.score(rank(validationSamples));
        }
    }

    @Override
    public Ranker createNew() {
        return new B200RFRanker();
    }
}
