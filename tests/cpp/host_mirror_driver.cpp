// host_mirror_driver.cpp — command-line driver of the C++ host mirror (ranklib_b200/host_cpp/ranklib_b200.hpp) for
// tests/test_zz_cpp_host_mirror.py.  Test infrastructure: every command prints plain text the Python test compares with
// the Python mirror / the reference's documented behaviour.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../ranklib_b200/host_cpp/ranklib_b200.hpp"

using namespace ranklib_b200;

static std::string slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

static int run(int argc, char** argv) {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "fmt") {  // lines "f <hex bits>" | "d <hex bits>" -> Float.toString / Double.toString
        std::string kind, hex;
        while (std::cin >> kind >> hex) {
            if (kind == "f") {
                const uint32_t b = (uint32_t)std::stoul(hex, nullptr, 16);
                float x;
                memcpy(&x, &b, 4);
                std::cout << javaFloatToString(x) << "\n";
            } else {
                const uint64_t b = std::stoull(hex, nullptr, 16);
                double x;
                memcpy(&x, &b, 8);
                std::cout << javaDoubleToString(x) << "\n";
            }
        }
        return 0;
    }
    if (cmd == "rand") {  // seed bound n
        JavaRandom r(std::stoll(argv[2]));
        const int bound = std::stoi(argv[3]), n = std::stoi(argv[4]);
        for (int i = 0; i < n; i++) std::cout << r.nextInt(bound) << (i + 1 < n ? " " : "\n");
        return 0;
    }
    if (cmd == "roundtrip") {  // model file -> loadFromString -> model()
        std::unique_ptr<Ranker> r = RankerFactory().createRanker(std::string(argv[3]) == "rf" ? RANKER_TYPE::RANDOM_FOREST : RANKER_TYPE::LAMBDAMART);
        r->loadFromString(slurp(argv[2]));
        std::cout << "FEATURES";
        for (int32_t f : r->getFeatures()) std::cout << " " << f;
        std::cout << "\n" << r->toString();
        return 0;
    }
    if (cmd == "read") {  // letor file [mustHaveRelDoc]
        const RankLists rl = FeatureManager::readInput(argv[2], argc > 3 && std::string(argv[3]) == "1");
        std::cout << rl.N << " " << rl.size() << " " << rl.F << "\n";
        for (int q = 0; q < rl.size(); q++) std::cout << rl.qids[(size_t)q] << ":" << rl.size(q) << (q + 1 < rl.size() ? " " : "\n");
        uint64_t h = 1469598103934665603ULL;  // FNV-1a over the raw bits of X (NaN canonical) and the labels
        auto mix = [&](uint32_t v) {
            for (int i = 0; i < 4; i++) {
                h ^= (v >> (8 * i)) & 0xff;
                h *= 1099511628211ULL;
            }
        };
        for (float x : rl.X) {
            uint32_t b;
            if (x != x) b = 0x7fc00000u; else memcpy(&b, &x, 4);
            mix(b);
        }
        for (float x : rl.label) {
            uint32_t b;
            memcpy(&b, &x, 4);
            mix(b);
        }
        std::cout << h << "\n";
        return 0;
    }
    if (cmd == "factory") {  // the rankers outside the accelerated path are refused with a RankLibError
        RankerFactory().createRanker(RANKER_TYPE::RANKNET);
        return 0;
    }
    if (cmd == "train") {  // letor ranker(0|6|8) metric nTrees nLeaves device [validation letor]
        auto train = std::make_shared<const RankLists>(FeatureManager::readInput(argv[2]));
        const int rt = std::stoi(argv[3]);
        const MetricScorer scorer = MetricScorerFactory().createScorer(argv[4]);
        const int device = std::stoi(argv[7]);
        std::shared_ptr<const RankLists> vali;
        if (argc > 8) vali = std::make_shared<const RankLists>(FeatureManager::readInput(argv[8]));
        if (rt == 8) {
            RFRanker::nBag = std::stoi(argv[5]);
            RFRanker::nTreeLeaves = std::stoi(argv[6]);
            RFRanker::seed = 11;
        } else {
            LambdaMART::nTrees = std::stoi(argv[5]);
            LambdaMART::nTreeLeaves = std::stoi(argv[6]);
            LambdaMART::nRoundToStopEarly = 3;
        }
        RankerTrainer trainer;
        std::unique_ptr<Ranker> r = trainer.train((RANKER_TYPE)rt, train, vali, {}, scorer, device);
        printf("NAME %s\nTRAIN %.17g\nVALI %.17g\n", r->name().c_str(), r->getScoreOnTrainingData(), r->getScoreOnValidationData());
        const auto ranks = r->rank(*train);
        printf("RANK0");
        for (int i : ranks[0]) printf(" %d", i);
        printf("\nMODEL\n%s", r->model().c_str());
        return 0;
    }
    std::cerr << "unknown command\n";
    return 2;
}

int main(int argc, char** argv) {
    try {
        return run(argc, argv);
    } catch (const RankLibError& e) {
        std::cout << "RankLibError: " << e.what() << "\n";
        return 3;
    }
}
