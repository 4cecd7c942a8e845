/*
 * B200MART — MART (R/learning/tree/MART.java) on the native path: the same context with kind = RLB_KIND_MART, which
 * selects pseudoResponses = label - modelScore (MART.java:47-51) and leaf output = float-chain mean (MART.java:54-65).
 * NOT COMPILED IN THIS IMAGE (no JDK).
 */
package ciir.umass.edu.learning.tree;

import java.util.List;

import ciir.umass.edu.learning.RankList;
import ciir.umass.edu.learning.Ranker;
import ciir.umass.edu.metric.MetricScorer;

public class B200MART extends B200LambdaMART {
    public B200MART() {
    }

    public B200MART(final List<RankList> samples, final int[] features, final MetricScorer scorer) {
        super(samples, features, scorer);
    }

    @Override
    protected int kind() {
        return NativeBridge.KIND_MART;
    }

    @Override
    public Ranker createNew() {
        return new B200MART();
    }

    @Override
    public String name() {
        return "MART";
    }
}
