"""CPU tests of the one-launch kernel's CTA layout (dzahui_b200/csrc/ol_layout.h, pure host code compiled here with g++): the
permutation the launcher derives from the probed SM placement must be a permutation whatever the placement looks like, and on
the placement a B200 actually produces (tests/golden/b200_cta_placement.json, recorded with DZB_OL_TRACE) it must give every
SM 33 of the 4883 tiles of BASELINE config 4's bar, the longer range to the CTA placed first."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def layout(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ol_layout") / "ol_layout_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-Wall", "-Wextra", "-o", exe, os.path.join(ROOT, "tests", "cpp", "ol_layout_check.cpp")])

    def run(smid):
        text = " ".join(str(v) for v in [len(smid)] + list(smid))
        out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
        return None if out[0] == "none" else np.array([int(v) for v in out[1:]])
    return run


def tile_ranges(perm, tiles, ragged):
    """the kernel's split (onelaunch_kernels.cuh: tb0, tb1) for positions perm[b]"""
    grid = len(perm)
    units = tiles + (1 if ragged else 0)
    tb0 = np.minimum(tiles, perm * units // grid)
    tb1 = np.where(perm + 1 == grid, tiles, np.minimum(tiles, (perm + 1) * units // grid))
    return tb0, tb1


def test_layout_on_the_recorded_b200_placement(layout):
    with open(os.path.join(ROOT, "tests", "golden", "b200_cta_placement.json")) as f:
        smid = np.array(json.load(f)["smid"])
    grid = len(smid)
    assert grid == 296 and len(set(smid.tolist())) == 148
    perm = layout(smid)
    assert perm is not None and sorted(perm.tolist()) == list(range(grid))
    tb0, tb1 = tile_ranges(perm, tiles=4883, ragged=True)     # 10^7 rows: 4882 full tiles + a ragged one
    count = tb1 - tb0
    assert count.sum() == 4883 and set(count.tolist()) <= {16, 17}
    order = np.argsort(perm)
    assert np.array_equal(tb0[order][1:], tb1[order][:-1]) and tb0[order][0] == 0 and tb1[order][-1] == 4883  # contiguous cover
    per_sm = {}
    for b in range(grid):
        per_sm.setdefault(int(smid[b]), []).append(b)
    totals = [sum(int(count[b]) for b in bs) for bs in per_sm.values()]
    assert max(totals) - min(totals) <= 1 and max(totals) == 33, sorted(set(totals))
    for bs in per_sm.values():
        first, second = sorted(bs)
        assert perm[first] == perm[second] + 1 and perm[second] % 2 == 0       # neighbours; the CTA placed first holds the odd position
        assert count[first] >= count[second]                                   # ... and with it the longer range
    # laid out by blockIdx the same placement gives half the SMs 32 tiles and half 34: what the probe is there to avoid
    c0 = np.subtract(*tile_ranges(np.arange(grid), 4883, True)[::-1])
    plain = sorted(set(sum(int(c0[b]) for b in bs) for bs in per_sm.values()))
    assert plain == [32, 33, 34]


def test_layout_is_a_permutation_for_any_uniform_placement(layout):
    rng = np.random.default_rng(11)
    for sms, per in [(148, 2), (7, 3), (1, 2), (132, 2), (5, 4)]:
        ids = rng.permutation(1000)[:sms]                 # sparse SM ids
        smid = rng.permutation(np.repeat(ids, per))
        perm = layout(smid)
        assert perm is not None and sorted(perm.tolist()) == list(range(sms * per))
        for s in ids:
            bs = np.sort(np.nonzero(smid == s)[0])
            pos = perm[bs]
            assert pos.max() - pos.min() == per - 1 and np.array_equal(pos, pos[0] - np.arange(per))  # consecutive, first CTA last


def test_layout_declines_what_it_does_not_understand(layout):
    assert layout([0, 1, 2, 3]) is None                   # one CTA per SM: nothing to pair
    assert layout([0, 0, 1]) is None                      # unequal shares
    assert layout([0, 0, 1, 1, 1, 2]) is None
    assert layout([0, 0, -1, -1]) is None                 # an SM id the probe did not fill in
    assert layout([]) is None
