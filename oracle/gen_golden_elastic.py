"""TEST INFRASTRUCTURE — tests/golden/elastic.npz from the UNMODIFIED reference `elastic_deform`
(skoots/train/merged_transform.py:75-188), run eagerly (PYTORCH_JIT=0) like gen_golden.py.

    PYTORCH_JIT=0 python oracle/gen_golden_elastic.py

PARITY STATUS.  The volume half (trilinear displacement field, identity grid, nearest grid_sample) is pinned: the
reference runs and the oracle restatement is asserted bit-equal to it below.  The skeleton half is UNPINNED: under this
image's torch 2.11 the reference's `_elastic_on_skeletons` (:43-72) cannot execute in any mode — its
`skel[ind, :] = grid[...]` is rejected ("Index put requires the source and destination dtypes match": the points are
integer tensors because they index the grid, the grid is fp32), scripted or eager.  So the reference is called with an
empty skeleton dict, and the skeleton outputs stored in the fixture come from the oracle's restatement of the intended
semantics (integer points take the grid's values by truncation toward zero, which is what the implicit cast of the
torch versions the reference was written for did).
"""
import os
import sys

os.environ["PYTORCH_JIT"] = "0"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import numpy as np
import torch

import ref_shim

ref_shim.install()
import skoots.train.merged_transform as mt  # noqa: E402
import skoots_oracle as orc  # noqa: E402


def main():
    pack = {}
    for tag, shape, ds, mag in (("a", (1, 1, 30, 26, 12), (6, 6, 2), (0.05, 0.05, 0.01)),
                                ("b", (1, 2, 25, 23, 14), (5, 4, 3), (0.2, 0.1, 0.05))):
        g = torch.Generator().manual_seed(len(tag) + shape[2])
        img = torch.rand(shape, generator=g)
        mask = ((torch.rand(shape, generator=g) > 0.6) * torch.randint(1, 9, shape, generator=g)).float()
        x, y, z = shape[2:]
        sk = {1: torch.stack([torch.randint(0, n, (40,), generator=g) for n in (x, y, z)], 1),
              2: torch.tensor([[x - 1, y - 1, z - 1], [0, 0, 0], [x + 1, 2, 3], [-1, 4, 4], [x // 2, y // 2, z // 2]])}
        torch.manual_seed(11)
        got = mt.elastic_deform(img, mask, skeleton={}, displacement_shape=ds, displacement_magnitude=mag)
        torch.manual_seed(11)
        noise = torch.rand((1, 3, ds[2], ds[1], ds[0]))
        mine = orc.elastic_deform(noise, img, mask, skeleton=sk, displacement_magnitude=mag)
        assert all(torch.equal(a, b) for a, b in zip(got[:2], mine[:2])), "oracle restatement differs from the reference (volumes)"
        pack.update({f"{tag}_image": img.numpy(), f"{tag}_mask": mask.numpy(), f"{tag}_noise": noise.numpy(),
                     f"{tag}_ds": np.array(ds), f"{tag}_mag": np.array(mag, dtype=np.float32),
                     f"{tag}_sk1": sk[1].numpy(), f"{tag}_sk2": sk[2].numpy(),
                     f"{tag}_out_image": got[0].numpy(), f"{tag}_out_mask": got[1].numpy(),
                     f"{tag}_out_sk1": mine[2][1].numpy(), f"{tag}_out_sk2": mine[2][2].numpy()})  # oracle (unpinned)
    path = os.path.join(ROOT, "tests", "golden", "elastic.npz")
    np.savez_compressed(path, **pack)
    print(f"elastic: {os.path.getsize(path) / 1024:.1f} KiB; oracle restatement == reference on the volumes of both cases")


if __name__ == "__main__":
    main()
