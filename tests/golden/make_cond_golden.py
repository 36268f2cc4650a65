"""Golden vectors of the conditioning producers (landmark maps, motion field, key-point maps, photo matting).

Run in the build container only (needs /root/reference):   python tests/golden/make_cond_golden.py

  * `draw2` and `cal_motion256` are the UNMODIFIED reference functions, imported from
    /root/reference/Module2/data/umlvdfw_test_dataset.py:34-81 (cwd = Module2, the module loads faceLmarkLookup.npy).
  * `kp_to_map` / `kp_to_map_some` (Module2/models/geomcgt_ifw_test_model.py:12-44): that module imports tensorflow at
    the top (photo2cartoon), which is absent, so the two function definitions are cut out of the reference file with
    `ast` and executed verbatim here -- the text is never copied into this repo.
  * The matting line is inline code of `GeomCGTIFWTestModel.forward` (geomcgt_ifw_test_model.py:280,292); it is
    executed here with the same torch ops.
The script asserts that oracle/cond_oracle.py reproduces all four and stores inputs + outputs in cond_*.npz.
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cond_oracle as O  # noqa: E402

REF = "/root/reference/Module2"


def reference_dataset_functions():
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    try:
        import data.umlvdfw_test_dataset as d
    finally:
        os.chdir(cwd)
    return d.draw2, d.cal_motion256


def reference_kp_functions():
    path = os.path.join(REF, "models", "geomcgt_ifw_test_model.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("kp_to_map", "kp_to_map_some"):
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns["kp_to_map_some"]


def edge_landmarks(rng):
    """Landmark sets that exercise clipping, half-way rounding and points outside the window."""
    lm = rng.uniform(-6, 262, (68, 2)).astype(np.float32)
    lm[:8] = [[0.5, 1.5], [2.5, 3.5], [255.5, 254.5], [-0.5, 10.5], [0, 0], [255, 255], [-3, 128], [258.4, 128]]
    return lm


def main():
    draw2, cal_motion256 = reference_dataset_functions()
    kp_to_map_some = reference_kp_functions()
    rng = np.random.RandomState(20261017)

    # ---- landmark maps (draw2 op 0), radius 3 as for crop_size 256 (umlvdfw_test_dataset.py:145)
    src, seq = O.landmark_sequence(6, seed=11)
    lands = np.concatenate([seq, edge_landmarks(rng)[None], edge_landmarks(rng)[None]]).astype(np.float32)
    ref = np.stack([draw2(256, 256, lands[t].copy(), 3, 2, op=0).numpy() for t in range(len(lands))])
    assert np.array_equal(O.draw_landmarks(lands), ref)
    ref5 = np.stack([draw2(512, 512, 2 * lands[t].copy(), 5, 4, op=0).numpy() for t in range(2)])
    assert np.array_equal(O.draw_landmarks(2 * lands[:2], 512, 5), ref5)
    np.savez_compressed(os.path.join(HERE, "cond_draw.npz"), lands=lands, bits=np.packbits(ref > 0))
    print("cond_draw", ref.shape, int((ref > 0).sum()))

    # ---- motion field (cal_motion256): smooth clip frames + scattered landmarks, some outside the window
    srcs = [src] * 6
    dsts = [seq[t] for t in range(6)]
    for _ in range(2):
        a = rng.uniform(-5, 260, (68, 2)).astype(np.float32)
        srcs.append(a)
        dsts.append((a + rng.normal(0, 6, (68, 2))).astype(np.float32))
    srcs, dsts = np.stack(srcs), np.stack(dsts)
    ref = np.stack([cal_motion256(torch.from_numpy(srcs[t].copy()), torch.from_numpy(dsts[t].copy()))
                    for t in range(len(dsts))]).astype(np.float32)
    mine = O.cal_motion_batch(srcs, dsts)
    err = float(np.abs(mine - ref).max())
    assert err <= 1e-6, err
    full = [0, 6]                                             # two frames in full, the rest on a 4x4-strided lattice
    np.savez_compressed(os.path.join(HERE, "cond_motion.npz"), src=srcs, dst=dsts, full_index=np.array(full),
                        full=ref[full], strided=ref[:, ::4, ::4])
    print("cond_motion", ref.shape, "oracle-vs-reference max-abs", err)

    # ---- key-point maps for netF: lm * 7/8 in float32 as flow_network_warp does (geomcgt_ifw_test_model.py:62-63)
    kps = np.stack([seq[0], edge_landmarks(rng)]).astype(np.float32)
    kps[1, 10] = [-8.0 / 7.0, 100.0]                          # becomes x == -1 after *7/8: the "missing point" branch
    kps78 = kps * 7 / 8
    assert kps78.dtype == np.float32 and kps78[1, 10, 0] == -1
    ref = kp_to_map_some((224, 224), kps78).numpy()
    assert ref.shape == (2, 68, 224, 224) and ref[1, 10].sum() == 0
    assert np.array_equal(O.kp_to_map(kps78), ref)
    np.savez_compressed(os.path.join(HERE, "cond_kp.npz"), kps=kps78, bits=np.packbits(ref > 0))
    print("cond_kp", ref.shape, int(ref.sum()))

    # ---- photo matting
    g = torch.Generator().manual_seed(77)
    real_A = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    matte = torch.rand(2, 1, 64, 64, generator=g)
    matte[0, 0, 0, :4] = torch.tensor([0.5, 0.50000006, 0.49999997, 1.0])
    mask = (matte > 0.5).float()                              # geomcgt_ifw_test_model.py:280
    out = ((real_A / 2 + 0.5) * mask + 1 - mask) * 2 - 1       # geomcgt_ifw_test_model.py:292
    o_out, o_mask = O.matte_photo(real_A.numpy(), matte.numpy())
    assert np.array_equal(o_out, out.numpy()) and np.array_equal(o_mask, mask.numpy())
    np.savez_compressed(os.path.join(HERE, "cond_matte.npz"), real_A=real_A.numpy(), matte=matte.numpy(),
                        out=out.numpy(), mask=mask.numpy())
    print("cond_matte", tuple(out.shape))


if __name__ == "__main__":
    main()
