"""The remap kernel's TMA form (ddn_align_config.use_tma: the depth tile arrives through one tensor copy) against the
reference's golden vectors and, bit for bit, against the default form.  Measured slower (profiles/r02_k3_tma_check.log),
so it is an option, not the default; the tests keep it honest."""

import numpy as np
import pytest
import torch

from depthdensifier_b200.synthetic import SceneConfig, make_scene
from oracle import restatement as R

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # refined depth within 1e-5 relative (north_star)


def _kmat(sc):
    V = sc.mono_depth.shape[0]
    kmat = torch.zeros((V, 3, 3), dtype=torch.float64, device=sc.mono_depth.device)
    kmat[:, 0, 0], kmat[:, 1, 1], kmat[:, 0, 2], kmat[:, 1, 2], kmat[:, 2, 2] = (sc.intrinsics[:, 0], sc.intrinsics[:, 1],
                                                                                 sc.intrinsics[:, 2], sc.intrinsics[:, 3], 1.0)
    return kmat


@pytest.mark.parametrize("V,W,H", [(5, 160, 120), (4, 200, 152), (3, 512, 384), (2, 132, 34)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_tma_form_is_bit_identical(lib_built, V, W, H, with_mask):
    """Partial tiles on the right and bottom edge, tiles whose halo starts left of / above the image."""
    from depthdensifier_b200 import ops

    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=600, seed=3), device=torch.device("cuda", 0))
    kmat = _kmat(sc)
    out = []
    for use_tma in (False, True):
        opts = ops.AlignOptions(zero_unmasked_passthrough=True, use_tma=use_tma)
        refined, stats = ops.align_views(sc.mono_depth, sc.mask if with_mask else None, sc.cam_from_world, kmat, sc.sparse_xyz,
                                         sc.sparse_offsets, 600, opts)
        torch.cuda.synchronize()
        out.append((refined.clone(), stats.clone()))
    assert torch.equal(out[0][1], out[1][1])
    assert torch.equal(out[0][0].view(torch.int32), out[1][0].view(torch.int32))
    if H >= 100:
        assert int((ops.decode_stats(out[0][1])[0]["status"])) == 0  # the remap path, not a pass-through, was compared


def test_tma_form_matches_reference_refiner_golden(lib_built, golden_dir):
    """Same golden cases as test_gpu_parity.test_align_matches_reference_refiner_golden (DepthRefiner.refine_depth run by the
    reference itself), through the TMA form."""
    from depthdensifier_b200 import ops

    g = np.load(golden_dir / "ref_refiner_cases.npz")
    specs = {
        "no_subsample": (dict(adaptive_correspondences=False), True),
        "skip_smoothing": (dict(adaptive_correspondences=False, skip_smoothing=True), True),
        "mask_none": (dict(adaptive_correspondences=False), False),
    }

    def cuda(x):
        return torch.from_numpy(np.ascontiguousarray(x)).cuda().contiguous()

    depth, pose = cuda(g["mono_depth"]), cuda(g["cam_from_world"])
    kmat = cuda(np.stack([R.kmatrix(i) for i in g["intrinsics"]]))
    sparse, off = cuda(g["sparse_xyz"]), cuda(g["sparse_offsets"])
    maxc = int(np.diff(g["sparse_offsets"]).max())
    for name, (kw, use_mask) in specs.items():
        mask = cuda(g["mask"]) if use_mask else None
        refined, stats = ops.align_views(depth, mask, pose, kmat, sparse, off, maxc, ops.AlignOptions(use_tma=True, **kw))
        torch.cuda.synchronize()
        st = ops.decode_stats(stats)
        refined = refined.cpu().numpy()
        for v in range(depth.shape[0]):
            ref = g[f"{name}/{v}/refined"]
            assert st[v]["status"] == 0
            assert np.array_equal(refined[v] == 0, ref == 0), (name, v)
            np.testing.assert_allclose(refined[v], ref, rtol=RTOL, atol=0, err_msg=f"{name}/{v}")


@pytest.mark.parametrize("W,H", [(162, 120), (128, 120), (160, 32)])
def test_tma_form_rejects_unsupported_shapes(lib_built, W, H):
    """Row pitch not a multiple of 16 bytes, or an image smaller than one 132 x 34 box: a clear error before any launch."""
    from depthdensifier_b200 import _lib, ops

    sc = make_scene(SceneConfig(n_views=2, width=W, height=H, n_sparse=100, seed=1), device=torch.device("cuda", 0))
    with pytest.raises(_lib.DDNError, match="use_tma"):
        ops.align_views(sc.mono_depth, sc.mask, sc.cam_from_world, _kmat(sc), sc.sparse_xyz, sc.sparse_offsets, 100,
                        ops.AlignOptions(use_tma=True))
