/*
 * DumpGoldens — runs the UNMODIFIED reference LambdaMART / MART on a LETOR file and writes, per boosting iteration, the
 * quantities the CPU oracle (oracle/ranklib_oracle.cpp) must reproduce: pseudo responses, weights, the fitted tree,
 * model scores and the training metric.  This is the pin the oracle is still missing (DESIGN.md section 5: "parity
 * unpinned" — no JVM in the build image).  Wherever a JDK exists:
 *
 *     scripts/make_java_goldens.sh /path/to/ranklib        ->  tests/golden/c1_java.txt
 *     python -m pytest tests/test_java_goldens.py          (compares the oracle with that file; skipped while it is absent)
 *
 * It lives in the reference's package to reach the protected members of LambdaMART
 * (R/learning/tree/LambdaMART.java:45-58,331,398,442); it subclasses, it does not edit.  The loop below issues the same
 * calls in the same order as LambdaMART.learn (:180-226) and adds only the dump.
 * Numbers are written as raw IEEE bits (hex) so that no decimal formatting stands between the JVM and the comparison.
 */
package ciir.umass.edu.learning.tree;

import java.io.PrintWriter;
import java.util.List;

import ciir.umass.edu.features.FeatureManager;
import ciir.umass.edu.learning.RankList;
import ciir.umass.edu.metric.MetricScorer;
import ciir.umass.edu.metric.MetricScorerFactory;
import ciir.umass.edu.utilities.MyThreadPool;

public class DumpGoldens {
    /** LambdaMART with a dumping copy of the boosting loop. */
    static class Dumping extends LambdaMART {
        final boolean mart;

        Dumping(final List<RankList> samples, final int[] features, final MetricScorer scorer, final boolean mart) {
            super(samples, features, scorer);
            this.mart = mart;
        }

        @Override
        protected void computePseudoResponses() {
            if (!mart) {
                super.computePseudoResponses();
                return;
            }
            for (int i = 0; i < martSamples.length; i++) { // MART.computePseudoResponses (MART.java:47-51)
                pseudoResponses[i] = martSamples[i].getLabel() - modelScores[i];
            }
        }

        @Override
        protected void updateTreeOutput(final RegressionTree rt) {
            if (!mart) {
                super.updateTreeOutput(rt);
                return;
            }
            for (final Split s : rt.leaves()) { // MART.updateTreeOutput (MART.java:54-65)
                float s1 = 0.0F;
                final int[] idx = s.getSamples();
                for (final int k : idx) {
                    s1 += pseudoResponses[k];
                }
                s.setOutput(s1 / idx.length);
            }
        }

        static void doubles(final PrintWriter out, final String tag, final double[] v) {
            final StringBuilder b = new StringBuilder(tag);
            for (final double x : v) {
                b.append(' ').append(Long.toHexString(Double.doubleToLongBits(x)));
            }
            out.println(b);
        }

        void dumpThresholds(final PrintWriter out) {
            for (int f = 0; f < thresholds.length; f++) {
                final StringBuilder b = new StringBuilder("THRESHOLDS " + f);
                for (final float t : thresholds[f]) {
                    b.append(' ').append(Integer.toHexString(Float.floatToIntBits(t)));
                }
                out.println(b);
            }
        }

        void learnAndDump(final PrintWriter out, final int trees) {
            ensemble = new Ensemble();
            dumpThresholds(out);
            for (int m = 0; m < trees; m++) {
                out.println("ITER " + (m + 1));
                computePseudoResponses();
                doubles(out, "LAMBDA", pseudoResponses);
                doubles(out, "WEIGHT", weights);
                hist.update(pseudoResponses);
                final RegressionTree rt = new RegressionTree(nTreeLeaves, martSamples, pseudoResponses, hist, minLeafSupport);
                rt.fit();
                ensemble.add(rt, learningRate);
                updateTreeOutput(rt);
                // leaves in RegressionTree.leaves() order: size, output bits, and the sample ids of the leaf
                final List<Split> leaves = rt.leaves();
                out.println("LEAVES " + leaves.size());
                for (final Split s : leaves) {
                    final int[] idx = s.getSamples();
                    final StringBuilder b = new StringBuilder("LEAF " + idx.length + " "
                            + Integer.toHexString(Float.floatToIntBits((float) s.getOutput())));
                    for (final int k : idx) {
                        b.append(' ').append(k);
                        modelScores[k] += learningRate * s.getOutput();
                    }
                    out.println(b);
                }
                out.println("TREE_BEGIN");
                out.print(rt.toString());
                out.println("TREE_END");
                rt.clearSamples();
                doubles(out, "SCORE", modelScores);
                scoreOnTrainingData = computeModelScoreOnTraining();
                out.println("METRIC " + Integer.toHexString(Float.floatToIntBits((float) scoreOnTrainingData)));
            }
        }
    }

    /** args: letorFile outFile metric(e.g. NDCG@10) trees leaves threads [mart] */
    public static void main(final String[] args) throws Exception {
        final List<RankList> samples = FeatureManager.readInput(args[0]);
        final int[] features = FeatureManager.getFeatureFromSampleVector(samples);
        final MetricScorer scorer = new MetricScorerFactory().createScorer(args[2]);
        LambdaMART.nTrees = Integer.parseInt(args[3]);
        LambdaMART.nTreeLeaves = Integer.parseInt(args[4]);
        MyThreadPool.init(Integer.parseInt(args[5]));
        final boolean mart = args.length > 6 && args[6].equals("mart");
        final Dumping r = new Dumping(samples, features, scorer, mart);
        r.init();
        try (PrintWriter out = new PrintWriter(args[1], "UTF-8")) {
            out.println("GOLDEN ranklib " + (mart ? "MART" : "LambdaMART") + " " + args[2] + " trees=" + args[3] + " leaves=" + args[4]
                    + " threads=" + args[5] + " java=" + System.getProperty("java.version"));
            r.learnAndDump(out, LambdaMART.nTrees);
        }
        MyThreadPool.getInstance().shutdown();
    }
}
