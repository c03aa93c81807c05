"""CPU: the full-volume-vs-sample comparison bench.py reports as `parity.sample_vs_oracle` (oracle/sample_check.py):
accepts a correct full-volume result, including components cut by the sample box, and rejects corrupted ones."""
import torch

import sample_check
import skoots_oracle as orc
from skoots_b200.synthetic import make_tube_volume

SCALE = torch.tensor((60, 60, 12))


def _case(mode, hops):
    full = (320, 300, 96) if mode == "whole" else (560, 540, 100)
    tv = make_tube_volume(full, 260 if mode == "whole" else 700, seed=4, want_mask=False, want_skeleton_dict=False)
    tv.skeleton[100:103, 10:290, 20] = 1   # a long component that the sample box cuts
    kw = dict(crop=sample_check.EVAL_CROP, overlap=sample_check.EVAL_OVERLAP, out_dtype=torch.int16) if mode == "eval" else {}
    got = orc.postprocess(tv.skeleton, tv.vectors, SCALE, N=hops, **kw)
    ladder = [((200, 200, 64), (64, 64, 16))] if mode == "whole" else [((500, 500, 90), (100, 100, 30))]
    pl = sample_check.plan(full, mode, hops, 1e9, ladder=ladder)
    R, S = pl["R"], pl["S"]
    want, labels, secs, kind = sample_check.run_cpu(tv.skeleton[:R[0], :R[1], :R[2]].contiguous(),
                                                    tv.vectors[:, :R[0], :R[1], :R[2]].contiguous(), SCALE, hops, mode)
    return full, got, want, labels, S


def test_sample_check_accepts_the_full_volume_result_whole_mode():
    full, got, want, labels, S = _case("whole", 1)
    res = sample_check.compare(got[:S[0], :S[1], :S[2]], want, labels, S, full)
    assert res["ok"], res
    assert res["labelled_compared"] > 1000 and res["labels_compared"] > 5
    assert res["excluded_voxels_of_components_cut_by_the_sample_box"] > 0   # the long component was excluded, not mis-compared
    bad = got.clone()
    x, y, z = (bad[:S[0], :S[1], :S[2]] > 0).nonzero()[17].tolist()
    bad[x, y, z] = 0
    assert not sample_check.compare(bad[:S[0], :S[1], :S[2]], want, labels, S, full)["ok"]
    bad = got.clone()
    vals = torch.unique(bad[:S[0], :S[1], :S[2]])
    vals = vals[vals > 0]
    bad[bad == vals[1]] = vals[2]                                            # two components merged
    assert not sample_check.compare(bad[:S[0], :S[1], :S[2]], want, labels, S, full)["ok"]


def test_sample_check_accepts_the_full_volume_result_eval_mode():
    full, got, want, labels, S = _case("eval", 3)
    res = sample_check.compare(got[:S[0], :S[1], :S[2]], want, labels, S, full)
    assert res["ok"], res
    assert res["labelled_compared"] > 1000
