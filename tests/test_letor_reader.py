"""The native LETOR reader (csrc/rlb_letor.cpp, rlb_letor_*) against a line-by-line Python transliteration of
FeatureManager.readInput (R/features/FeatureManager.java:187-245) + DataPoint.parse (R/learning/DataPoint.java:58-110).

Host-only code: every test here runs without a GPU.  The checker below is test infrastructure (pure-Python loops,
small inputs); Float.parseFloat is modelled by exact rational rounding to the nearest float (_round_f32).
"""
import os
import re

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from ranklib_b200.host import native, rankers as R, synth

JAVA_WS = " \t\n\x0b\f\r"


def _java_trim(s):
    a, b = 0, len(s)
    while a < b and s[a] <= " ":
        a += 1
    while b > a and s[b - 1] <= " ":
        b -= 1
    return s[a:b]


_DEC = re.compile(r"^[+-]?(NaN|Infinity|((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?[fFdD]?))$")


def _round_f32(text):
    """The float nearest to the decimal string, ties to even — computed in exact rational arithmetic.  (numpy's
    np.float32(str) goes through a double and rounds twice: np.float32("1.0000000596046448") is 1.0, the JDK's
    Float.parseFloat and glibc's strtof give 1.0000001.)"""
    from decimal import Decimal
    from fractions import Fraction
    x = Fraction(Decimal(text))
    neg = text.strip().startswith("-")
    a = abs(x)
    if a == 0:
        return np.float32(-0.0 if neg else 0.0)
    # a = m * 2^e with 2^23 <= m < 2^24 (normal) or e = -149 (subnormal)
    e = a.numerator.bit_length() - a.denominator.bit_length() - 24
    while Fraction(2) ** (e + 24) <= a:
        e += 1
    while Fraction(2) ** (e + 23) > a:
        e -= 1
    e = max(e, -149)
    q = a / Fraction(2) ** e
    m = q.numerator // q.denominator
    r = q - m
    if r > Fraction(1, 2) or (r == Fraction(1, 2) and (m & 1)):
        m += 1
    v = np.float32(np.inf) if m * Fraction(2) ** e >= Fraction(2) ** 128 else np.float32(float(m) * 2.0 ** e)
    return np.float32(-v) if neg else v


def _java_parse_float(tok):
    """Float.parseFloat for the decimal forms (hex literals are tested separately)."""
    t = _java_trim(tok)
    if not _DEC.match(t):
        raise ValueError(tok)
    u = t.lstrip("+-")
    if u == "NaN":
        return np.float32(np.nan)
    if u == "Infinity":
        return np.float32(-np.inf if t.startswith("-") else np.inf)
    return _round_f32(t.rstrip("fFdD"))


def _ref_read(path, must_have_rel=False):
    """readInput + parse, written to follow the Java statement by statement."""
    lists, rl, last_id, has_rel, entries, maxf = [], [], "", False, 0, 0
    with open(path, newline="\n") as fh:
        for content in fh.read().split("\n"):
            content = _java_trim(content)
            if len(content) == 0 or content.find("#") == 0:
                continue
            text = content
            idx = text.find("#")
            if idx != -1:
                text = _java_trim(text[:idx])
            fs = [t for t in re.split("[" + JAVA_WS + "]+", text)]
            label = _java_parse_float(fs[0])
            if label < 0:
                raise ValueError("negative label")
            qid = fs[1][fs[1].rfind(":") + 1:]
            feats = {}
            for tok in fs[2:]:
                key = tok[:tok.index(":")]
                val = tok[tok.rfind(":") + 1:]
                if not re.match(r"^[+-]?\d+$", key):
                    raise ValueError("fid")
                f = int(key)
                if f <= 0:
                    raise ValueError("fid <= 0")
                feats[f] = _java_parse_float(val)
                maxf = max(maxf, f)
            if last_id != "" and last_id != qid:
                if not must_have_rel or has_rel:
                    lists.append(rl)
                rl, has_rel = [], False
            if label > 0:
                has_rel = True
            last_id = qid
            rl.append((label, qid, feats))
            entries += 1
    if rl and (not must_have_rel or has_rel):
        lists.append(rl)
    N = sum(len(r) for r in lists)
    X = np.full((N, maxf), np.nan, np.float32)
    label = np.zeros(N, np.float32)
    qoff, qids, i = [0], [], 0
    for r in lists:
        qids.append(r[0][1])
        for lab, _, feats in r:
            label[i] = lab
            for f, v in feats.items():
                X[i, f - 1] = v
            i += 1
        qoff.append(i)
    return X, label, np.array(qoff, np.int32), qids, entries


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)) if a.dtype == np.float32 else np.array_equal(a, b)


def _check(path, must_have_rel=False, nthreads=0):
    X, label, qoff, fids, qids, entries = native.read_letor(path, must_have_rel, None, nthreads)
    rX, rl, rq, rids, rent = _ref_read(path, must_have_rel)
    assert X.shape == rX.shape
    assert np.array_equal(np.isnan(X), np.isnan(rX))
    assert np.array_equal(np.nan_to_num(X).view(np.uint32), np.nan_to_num(rX).view(np.uint32))   # bit-exact incl. -0.0
    assert _same(label, rl) and np.array_equal(qoff, rq) and qids == rids and entries == rent
    assert np.array_equal(fids, np.arange(1, X.shape[1] + 1))
    return X, label, qoff, qids


TRICKY = """\
# a comment line, then a blank one

2 qid:a 1:0.5 2:1e-3 5:7 # doc-1 with: colons 9:9
0 qid:a 3:-0.0 1:+.25   2:3.   # trailing blanks and tabs
\t1\tqid:a\t1:1.0f\t2:2D 4:1e2f
0 qid:b 1:NaN 2:Infinity 3:-Infinity
3.5 qid:b 7:16777217 6:0.1 5:1.17549435E-38 4:3.4028235e38
0 qid:b 2:1.0 2:2.0 2:3.0
1 qid:c:d:e 1:x:y:0.75 10:4.9e-324 9:1e-46 8:1e39
0 c 1:00012.500 2:123456789012345678901234567890 3:0.000000000000000000000000000001
0 qid: 1:1
1 qid: 1:2
0 qid:z 1:8.5
1 qid:a 1:9.5
"""


def test_reader_follows_the_reference_line_rules(tmp_path):
    p = tmp_path / "tricky.txt"
    p.write_bytes(TRICKY.encode())
    X, label, qoff, qids = _check(str(p))
    # ids: after the LAST ':' ("qid:c:d:e" -> "e"; "c" -> "c"; "qid:" -> ""), a change closes the list, an EMPTY last id
    # never does (FeatureManager.java:215: the "z" line joins the list of the two empty ids, whose getID() is that of
    # its first point), and a repeated id later in the file opens a new list
    assert qids == ["a", "b", "e", "c", "", "a"]
    assert list(qoff) == [0, 3, 6, 7, 8, 11, 12]
    assert X[0, 4] == 7 and np.isnan(X[0, 2])                       # "# doc-1 with: colons 9:9" is description
    assert np.signbit(X[1, 2]) and X[1, 2] == 0                      # -0.0 kept
    assert X[2, 0] == 1 and X[2, 1] == 2 and X[2, 3] == 100          # f / D suffixes
    assert np.isnan(X[3, 0]) and np.isinf(X[3, 1]) and X[3, 2] == -np.inf
    assert X[4, 6] == np.float32(16777216.0)                         # 16777217 rounds to even
    assert X[5, 1] == 3                                               # a repeated fid overwrites
    assert X[6, 0] == np.float32(0.75)                               # fid before the FIRST ':', value after the LAST
    assert X[6, 9] == 0 and X[6, 8] == 0 and np.isinf(X[6, 7])      # underflow to 0, overflow to Infinity
    _check(str(p), must_have_rel=True)
    kept = native.read_letor(str(p), True)
    assert len(kept[1]) == 11 and kept[4] == ["a", "b", "e", "", "a"]   # list "c" (no relevant doc) is dropped
    assert kept[5] == 12                                              # entries read before the filter


def test_must_have_relevant_doc_filter(tmp_path):
    p = tmp_path / "f.txt"
    p.write_text("0 qid:1 1:1\n0 qid:1 1:2\n1 qid:2 1:3\n0 qid:3 1:4\n0 qid:4 1:5\n2 qid:4 1:6\n0 qid:5 1:7\n")
    X, label, qoff, fids, qids, entries = native.read_letor(str(p), True)
    assert qids == ["2", "4"] and list(qoff) == [0, 1, 3] and list(X[:, 0]) == [3, 5, 6] and entries == 7
    _check(str(p), True)
    _check(str(p), False)
    rl = R.read_letor(str(p), must_have_rel_doc=True)
    assert rl.size() == 2 and rl.qids == ["2", "4"]


@pytest.mark.parametrize("bad,msg", [
    ("-1 qid:1 1:0.5\n", "negative"),
    ("1 qid:1 0:0.5\n", "less than or equal to zero"),
    ("1 qid:1 -3:0.5\n", "less than or equal to zero"),
    ("1 qid:1 a:0.5\n", "feature id"),
    ("1 qid:1 1.0:0.5\n", "feature id"),
    ("1 qid:1 1:abc\n", "feature value"),
    ("1 qid:1 1:inf\n", "feature value"),          # Java has no "inf" / "nan" spellings
    ("1 qid:1 1:0x10\n", "feature value"),         # a hex literal needs its binary exponent
    ("1 qid:1 15\n", "without ':'"),
    ("x qid:1 1:1\n", "label"),
    ("1\n", "label and a qid"),
])
def test_malformed_lines_raise_ranklib_error(tmp_path, bad, msg):
    p = tmp_path / "bad.txt"
    p.write_text("0 qid:0 1:1\n" * 3 + bad)
    with pytest.raises(native.RankLibError, match=msg) as ei:
        native.read_letor(str(p))
    assert "line 4" in str(ei.value) and "readInput" in str(ei.value)
    with pytest.raises(native.RankLibError):
        native.read_letor(str(tmp_path / "does-not-exist.txt"))


def test_empty_and_comment_only_files(tmp_path):
    for body in ["", "\n\n", "# nothing\n   \n#x"]:
        p = tmp_path / "e.txt"
        p.write_text(body)
        X, label, qoff, fids, qids, entries = native.read_letor(str(p))
        assert X.shape == (0, 0) and len(label) == 0 and list(qoff) == [0] and qids == [] and entries == 0


def test_file_without_trailing_newline_and_crlf(tmp_path):
    p = tmp_path / "n.txt"
    p.write_bytes(b"1 qid:1 1:0.5 2:0.25\r\n0 qid:1 2:4\r\n2 qid:2 1:8")
    X, label, qoff, qids = _check(str(p))
    assert X.shape == (3, 2) and list(qoff) == [0, 2, 3] and X[2, 0] == 8 and np.isnan(X[2, 1])


@pytest.mark.parametrize("nthreads", [1, 2, 7, 0])
def test_multithreaded_slabs_equal_single_thread(tmp_path, nthreads):
    """A file large enough to be cut into many slabs: queries straddle the slab borders; selected / missing / repeated
    feature columns."""
    rng = np.random.default_rng(9)
    Q, F = 300, 23
    sizes = rng.integers(1, 40, Q)
    N = int(sizes.sum())
    X = rng.normal(size=(N, F)).astype(np.float32)
    X[rng.random((N, F)) < 0.3] = 0
    label = rng.integers(0, 5, N).astype(np.float32)
    qoff = np.zeros(Q + 1, np.int32)
    qoff[1:] = np.cumsum(sizes)
    p = tmp_path / "big.txt"
    synth.write_letor(str(p), X, label, qoff)
    assert os.path.getsize(p) > 64 * 4096
    Xn, ln, qn, fids, qids, entries = native.read_letor(str(p), False, None, nthreads)
    assert np.array_equal(Xn.view(np.uint32), X.view(np.uint32)) and np.array_equal(ln, label) and np.array_equal(qn, qoff)
    assert entries == N and len(qids) == Q
    sel = np.array([5, 1, 40, 5, 23], np.int32)        # out of order, beyond max fid, repeated
    Xs = native.read_letor(str(p), False, sel, nthreads)[0]
    assert np.array_equal(Xs[:, 0], X[:, 4]) and np.array_equal(Xs[:, 3], X[:, 4]) and np.array_equal(Xs[:, 1], X[:, 0])
    assert np.isnan(Xs[:, 2]).all() and np.array_equal(Xs[:, 4], X[:, 22])


def test_sparse_lines_leave_unknowns(tmp_path):
    rng = np.random.default_rng(2)
    lines, want = [], []
    for i in range(500):
        feats = {int(f): np.float32(rng.normal()) for f in rng.choice(np.arange(1, 60), rng.integers(0, 12), replace=False)}
        want.append(feats)
        toks = " ".join(f"{f}:{float(v)!r}" for f, v in feats.items())
        lines.append(f"{i % 3} qid:{i // 10} {toks} # d{i}")
    p = tmp_path / "s.txt"
    p.write_text("\n".join(lines) + "\n")
    X, label, qoff, qids = _check(str(p), nthreads=3)
    for i, feats in enumerate(want):
        row = np.full(X.shape[1], np.nan, np.float32)
        for f, v in feats.items():
            row[f - 1] = v
        assert np.array_equal(np.isnan(X[i]), np.isnan(row)) and np.array_equal(np.nan_to_num(X[i]), np.nan_to_num(row))


# ---- Float.parseFloat: correctly rounded decimal -> float, every path of the parser --------------------------------
KNOWN = [("0", 0.0), ("-0", -0.0), ("1", 1.0), ("0.1", 0.1), ("16777217", 16777216.0), ("16777219", 16777220.0),
         ("3.4028235e38", 3.4028235e38), ("3.4028236e38", np.inf), ("1.17549435E-38", 1.17549435e-38), ("1e-45", 1.4e-45),
         ("7e-46", 0.0), ("7.1e-46", 1.4e-45), ("4.9e-324", 0.0), ("1.4E-45", 1.4e-45), (".5", 0.5), ("5.", 5.0),
         ("1e5f", 1e5), ("2.5D", 2.5), ("+3", 3.0), ("NaN", np.nan), ("-Infinity", -np.inf), ("0x1.8p1", 3.0),
         ("0x1p-149", 1.4e-45), ("-0X.8P0f", -0.5), ("0x1.000001p0", 1.0), ("0x1.000003p0d", 1.0000002384185791),
         # float midpoints and their decimal neighbours: the double-rounding hazards of a double fast path
         ("1.00000005960464477539062500", 1.0), ("1.00000005960464477539062501", 1.0000001192092896),
         ("1.00000017881393432617187500", 1.0000002384185791), ("1.0000001788139343", 1.0000001192092896),
         ("1.0000001788139344", 1.0000002384185791),
         ("1.0000000596046448", 1.0000001192092896), ("1.0000000596046447", 1.0),
         ("8388608.5", 8388608.0), ("8388609.5", 8388610.0), ("0.000000000000000000000000000000000000011754942", 1.1754942e-38)]


def test_java_float_known_answers(built):
    for text, want in KNOWN:
        got = native.parse_java_float(text)
        assert got is not None, text
        w = np.float32(want)
        assert (np.isnan(got) and np.isnan(w)) or got.view(np.uint32) == w.view(np.uint32), (text, got, w)
        if _DEC.match(text) and text.lstrip("+-") not in ("NaN", "Infinity"):    # the hand-written table against exact rounding
            assert _round_f32(text.rstrip("fFdD")).view(np.uint32) == w.view(np.uint32), text
    for text in ["", " ", ".", "e5", "1e", "1e+", "inf", "nan", "NAN", "infinity", "0x10", "0x", "1,5", "1 2", "--1", "1ff", "1e5e5",
                 "0x1p1.5", "١"]:
        assert native.parse_java_float(text) is None, text
    assert native.parse_java_float("  \t1.5 \x01") == np.float32(1.5)   # String.trim drops every char <= ' '


@settings(max_examples=3000, deadline=None)
@given(st.integers(0, 2 ** 64 - 1), st.integers(-60, 50), st.booleans())
def test_java_float_matches_correct_rounding_on_random_decimals(mant, exp10, as_plain):
    """Random decimal strings (integer mantissas of any length up to 20 digits, exponents across the float range):
    the native parser must agree with numpy's correctly rounded conversion bit for bit."""
    text = f"{mant}e{exp10}"
    if as_plain and -30 <= exp10 <= 0:
        s = str(mant).rjust(-exp10 + 1, "0")
        text = s[:len(s) + exp10] + "." + s[len(s) + exp10:] if exp10 < 0 else s
    got = native.parse_java_float(text)
    want = _round_f32(text)
    assert got is not None and got.view(np.uint32) == want.view(np.uint32), (text, got, want)


@settings(max_examples=2000, deadline=None)
@given(st.integers(0, 2 ** 23 - 1), st.integers(-140, 120), st.integers(-3, 3), st.integers(9, 40))
def test_java_float_near_float_midpoints(frac, e2, delta, digits):
    """Decimal expansions (9..40 significant digits: below 16 the double fast path, above it strtof) of float midpoints and of their neighbours a few double-ulps
    away — exactly the inputs on which rounding twice (decimal -> double -> float) goes wrong."""
    from decimal import Decimal, getcontext
    getcontext().prec = 80
    lo = (Decimal(2 ** 23 + frac) * Decimal(2) ** (e2 - 23))
    mid = lo + Decimal(2) ** (e2 - 24)
    x = mid + Decimal(delta) * Decimal(2) ** (e2 - 53)
    text = f"{x:.{digits - 1}e}"
    got = native.parse_java_float(text)
    want = _round_f32(text)
    assert got is not None and got.view(np.uint32) == want.view(np.uint32), (text, got, want)


def test_gzip_input_and_lone_carriage_returns(tmp_path):
    """FileUtils.smartReader (R/utilities/FileUtils.java:40-46) reads *.gz through GZIPInputStream, and
    BufferedReader.readLine ends a line at \\n, \\r or \\r\\n."""
    import gzip
    X, label, qoff = synth.c1()
    p = tmp_path / "t.txt"
    synth.write_letor(str(p), X[:300, :12], label[:300], qoff[:8])
    raw = p.read_bytes()
    want = native.read_letor(str(p))
    (tmp_path / "t.txt.gz").write_bytes(gzip.compress(raw))
    got = native.read_letor(str(tmp_path / "t.txt.gz"))
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)) and np.array_equal(got[2], want[2]) and got[4] == want[4]
    (tmp_path / "cr.txt").write_bytes(raw.replace(b"\n", b"\r"))          # classic Mac line ends
    got = native.read_letor(str(tmp_path / "cr.txt"))
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)) and np.array_equal(got[2], want[2])
    (tmp_path / "mixed.txt").write_bytes(raw.replace(b"\n", b"\r\n", 100).replace(b"\n1 ", b"\r1 ", 50))
    got = native.read_letor(str(tmp_path / "mixed.txt"), False, None, 4)
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)) and np.array_equal(got[2], want[2])
    (tmp_path / "bad.gz").write_bytes(b"1 qid:1 1:2\n")                    # GZIPInputStream: "Not in GZIP format"
    with pytest.raises(native.RankLibError, match="Not in GZIP format"):
        native.read_letor(str(tmp_path / "bad.gz"))
    (tmp_path / "cut.gz").write_bytes(gzip.compress(raw)[:-20])            # truncated stream
    with pytest.raises(native.RankLibError):
        native.read_letor(str(tmp_path / "cut.gz"))
    (tmp_path / "err.txt").write_bytes(b"1 qid:1 1:1\r\n\r\n1 qid:1 1:x\r\n")
    with pytest.raises(native.RankLibError, match="line 3"):
        native.read_letor(str(tmp_path / "err.txt"))


def test_binary_cache_round_trip(tmp_path):
    p = tmp_path / "tricky.txt"
    p.write_bytes(TRICKY.encode())
    b = tmp_path / "tricky.rlb"
    native.letor_to_binary(str(p), str(b))
    for must in (False, True):
        t = native.read_letor(str(p), must)
        c = native.read_letor(str(b), must)
        assert np.array_equal(np.isnan(t[0]), np.isnan(c[0])) and np.array_equal(np.nan_to_num(t[0]).view(np.uint32), np.nan_to_num(c[0]).view(np.uint32))
        assert np.array_equal(t[1], c[1]) and np.array_equal(t[2], c[2]) and np.array_equal(t[3], c[3]) and t[4] == c[4] and t[5] == c[5]
    sel = np.array([3, 99, 1], np.int32)
    assert np.array_equal(np.nan_to_num(native.read_letor(str(b), False, sel)[0]), np.nan_to_num(native.read_letor(str(p), False, sel)[0]))
    # a larger set through several fill threads, and a cache of a cache
    X, label, qoff = synth.c1()
    big = tmp_path / "big.txt"
    synth.write_letor(str(big), X, label, qoff)
    native.letor_to_binary(str(big), str(tmp_path / "big.rlb"))
    native.letor_to_binary(str(tmp_path / "big.rlb"), str(tmp_path / "big2.rlb"))
    assert (tmp_path / "big.rlb").read_bytes() == (tmp_path / "big2.rlb").read_bytes()
    got = native.read_letor(str(tmp_path / "big2.rlb"))
    assert np.array_equal(got[0].view(np.uint32), X.view(np.uint32)) and np.array_equal(got[1], label) and np.array_equal(got[2], qoff)
    # damaged caches are refused, not read past their end
    raw = (tmp_path / "big.rlb").read_bytes()
    for cut in (20, 60, 5000, len(raw) - 4):
        (tmp_path / "cut.rlb").write_bytes(raw[:cut])
        with pytest.raises(native.RankLibError, match="binary cache"):
            native.read_letor(str(tmp_path / "cut.rlb"))
    bad = bytearray(raw)
    bad[16:24] = (2 ** 62).to_bytes(8, "little")
    (tmp_path / "bad.rlb").write_bytes(bytes(bad))
    with pytest.raises(native.RankLibError, match="binary cache"):
        native.read_letor(str(tmp_path / "bad.rlb"))
