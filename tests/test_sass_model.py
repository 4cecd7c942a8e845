"""No GPU needed: the SASS of the histogram kernels, read with scripts/sass_stalls.py.

The consumers of k_hist_root / k_hist_child run one warp per scheduler, so the sum of the stall counts ptxas wrote into the
control words of the stage loop is the loop's issue time.  These tests keep what that model found: the
default kernels' stage loops carry no block of register moves (the round-2 kernels paid 25 per stage for a branch inside
the loop) and need fewer issue cycles than the kernels they replaced."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
OBJ = os.path.join(ROOT, "ranklib_b200", "csrc", "rlb_boost.o")


def _opcode(text):
    p = text.split()
    return p[1] if p[0].startswith("@") else p[0]


def _stage_loop(pattern, marker, count):
    """the loop of the kernel that holds exactly `count` instructions whose opcode starts with `marker`"""
    import re

    import sass_stalls as ss
    ins = ss.parse(OBJ, pattern)
    assert ins, pattern
    index = {d["addr"]: i for i, d in enumerate(ins)}
    for i, d in enumerate(ins):
        m = re.search(r"BRA(?:\.[A-Z.]+)?\s+(0x[0-9a-f]+)", d["text"])
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= d["addr"] and tgt in index:
            body = ins[index[tgt]:i + 1]
            if sum(_opcode(x["text"]).startswith(marker) for x in body) == count:
                return body
    raise AssertionError(f"no loop with {count} x {marker} in {pattern}")


def _cost(body):
    return sum(max(x["stall"], 1) for x in body)


def _moves(body):
    return sum("IMAD.MOV" in x["text"] or x["text"].split()[0] == "MOV" for x in body)


@pytest.fixture(scope="module")
def objfile(built):
    if not os.path.exists(OBJ):
        pytest.skip("rlb_boost.o not kept by this build")
    return OBJ


def test_root_stage_loop_has_no_register_move_block(objfile):
    new = _stage_loop("k_hist_rootILi1E", "LDS.64", 32)      # 4 chunks x 8 read-modify-writes per stage
    old = _stage_loop("k_hist_rootILi0E", "LDS.64", 32)
    assert _moves(old) >= 20, "the round-2 kernel is the one with the move block (if not, the comparison is void)"
    assert _moves(new) <= 4, [x["text"] for x in new if "MOV" in x["text"]]
    assert len(new) < len(old) and _cost(new) < _cost(old), (len(new), len(old), _cost(new), _cost(old))


def test_child_stage_loop_keeps_what_the_variants_bought(objfile):
    new = _stage_loop("k_hist_childILi1E", "LDS.U16", 32)    # 4 chunks x 8 bins per stage
    old = _stage_loop("k_hist_childILi0E", "LDS.U16", 32)
    ops = lambda body, name: sum(_opcode(x["text"]).startswith(name) for x in body)
    assert ops(new, "LDS.128") >= 16 and ops(old, "LDS.128") == 0      # a chunk's responses as four 16-byte loads
    assert ops(new, "IMAD.IADD") <= 4 and ops(old, "IMAD.IADD") >= 20  # column offset folded into the address multiply-add
    assert _moves(old) >= 20 and _moves(new) <= 10                      # no register-move block at the stage boundary
