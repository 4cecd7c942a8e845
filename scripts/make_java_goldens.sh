#!/usr/bin/env bash
# make_java_goldens.sh — pins the CPU oracle against the REAL reference, wherever a JDK exists (there is none in the
# build image: DESIGN.md section 5, "parity unpinned").
#
#   scripts/make_java_goldens.sh /path/to/codelibs-ranklib-checkout [trees] [threads]
#
# Compiles the reference's own sources (only what LambdaMART needs, straight from the checkout: no Maven, nothing is
# copied into this repository) together with scripts/java/.../DumpGoldens.java into a scratch directory, writes the
# seeded C1 workload as LETOR text, runs the unmodified LambdaMART and MART on it and stores what they computed in
#   tests/golden/c1_java.txt  and  tests/golden/c1_java_mart.txt
# `python -m pytest tests/test_java_goldens.py` then compares the oracle with those files (the test is skipped while
# they are absent).  Commit the two files: from then on the oracle is pinned by the reference itself.
set -euo pipefail
REF=${1:?path to a checkout of codelibs/ranklib}
TREES=${2:-20}
THREADS=${3:-1}
HERE=$(cd "$(dirname "$0")/.." && pwd)
command -v javac >/dev/null || { echo "javac not found: this script needs a JDK (8 or newer)"; exit 2; }
WORK=$(mktemp -d)
trap 'rm -rf "$WORK"' EXIT
SRC="$REF/src/main/java"
javac -nowarn -d "$WORK/classes" -sourcepath "$SRC:$HERE/scripts/java" \
      "$HERE/scripts/java/ciir/umass/edu/learning/tree/DumpGoldens.java"
python - "$WORK/c1.txt" <<PY
import sys
sys.path.insert(0, "$HERE")
from ranklib_b200.host import synth
X, label, qoff = synth.c1()
synth.write_letor(sys.argv[1], X, label, qoff)
PY
mkdir -p "$HERE/tests/golden"
java -cp "$WORK/classes" ciir.umass.edu.learning.tree.DumpGoldens "$WORK/c1.txt" "$HERE/tests/golden/c1_java.txt" NDCG@10 "$TREES" 10 "$THREADS"
java -cp "$WORK/classes" ciir.umass.edu.learning.tree.DumpGoldens "$WORK/c1.txt" "$HERE/tests/golden/c1_java_mart.txt" NDCG@10 "$TREES" 10 "$THREADS" mart
echo "wrote tests/golden/c1_java.txt and tests/golden/c1_java_mart.txt; now run: python -m pytest tests/test_java_goldens.py"
