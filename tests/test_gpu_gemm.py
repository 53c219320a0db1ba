"""GPU tests of the DMMA/TMA GEMM engine on its own: every tile configuration, both A layouts, ragged
shapes, batch strides and shared operands, against torch.matmul in FP64 (the one floating-point kernel
for which a torch reference is the right oracle)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_CFGS = 12


def _run(torch, M, N, K, batch, a_mmajor, shared_a, cfg):
    from hermipy_b200.plan import dgemm
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + batch)
    ev = lambda v: (v + 1) & ~1
    nb_a = 1 if shared_a else batch
    if a_mmajor:
        A = torch.randn((nb_a, K, ev(M)), generator=g, dtype=torch.float64, device="cuda")
        A_log = A[:, :, :M].transpose(1, 2)
        lda = ev(M)
    else:
        A = torch.randn((nb_a, M, ev(K)), generator=g, dtype=torch.float64, device="cuda")
        A_log = A[:, :, :K]
        lda = ev(K)
    B = torch.randn((batch, K, ev(N)), generator=g, dtype=torch.float64, device="cuda")
    C = torch.full((batch, M, ev(N)), float("nan"), dtype=torch.float64, device="cuda")
    os.environ["HB_GEMM_CFG"] = str(cfg)
    try:
        dgemm(M, N, K, A, lda, B, ev(N), C, ev(N), batch=batch, strideA=0 if shared_a else A.stride(0),
              strideB=B.stride(0), strideC=C.stride(0), a_mmajor=a_mmajor)
    finally:
        os.environ["HB_GEMM_CFG"] = "-1"
    want = torch.matmul(A_log, B[:, :, :N])
    got = C[:, :, :N]
    assert torch.isfinite(got).all(), "unwritten output entries"
    scale = torch.matmul(A_log.abs(), B[:, :, :N].abs()).max().item()
    return (got - want).abs().max().item() / scale


@pytest.mark.parametrize("cfg", list(range(N_CFGS)) + [-1])
@pytest.mark.parametrize("a_mmajor", [False, True])
def test_every_tile_configuration(cfg, a_mmajor):
    import torch
    for (M, N, K, batch, shared) in [(300, 200, 100, 1, False), (33, 209, 17, 3, True), (257, 131, 402, 2, False),
                                     (32, 256, 256, 4, True), (1, 64, 64, 1, False), (515, 333, 77, 2, True)]:
        err = _run(torch, M, N, K, batch, a_mmajor, shared, cfg)
        assert err < 1e-14, (cfg, a_mmajor, M, N, K, batch, shared, err)


def test_tile_configurations_agree_bitwise():
    """Every configuration accumulates each output over k in the same order: the autotuner's choice (which depends
    on timing) must never change a result."""
    import torch
    from hermipy_b200.plan import dgemm
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K, batch = 530, 402, 201, 3
    A = torch.randn((M, K + 1), generator=g, dtype=torch.float64, device="cuda")
    B = torch.randn((batch, K, N), generator=g, dtype=torch.float64, device="cuda")
    results = []
    for cfg in range(N_CFGS):
        C = torch.zeros((batch, M, N), dtype=torch.float64, device="cuda")
        os.environ["HB_GEMM_CFG"] = str(cfg)
        try:
            dgemm(M, N, K, A, K + 1, B, N, C, N, batch=batch, strideA=0, strideB=B.stride(0), strideC=C.stride(0))
        finally:
            os.environ["HB_GEMM_CFG"] = "-1"
        results.append(C)
    for cfg in range(1, N_CFGS):
        assert torch.equal(results[cfg], results[0]), cfg


def test_alignment_is_checked():
    import torch
    from hermipy_b200 import HermiteB200Error
    from hermipy_b200.plan import dgemm
    A = torch.zeros((8, 7), dtype=torch.float64, device="cuda")     # odd leading dimension
    B = torch.zeros((7, 8), dtype=torch.float64, device="cuda")
    C = torch.zeros((8, 8), dtype=torch.float64, device="cuda")
    with pytest.raises(HermiteB200Error):
        dgemm(8, 8, 7, A, 7, B, 8, C, 8)


@pytest.mark.parametrize("cfg", list(range(N_CFGS)))
def test_stage_release_waits_for_fragment_loads(cfg):
    """Regression: short K (few k-tiles per tile), many tiles per CTA, repeated launches.  A stage of the TMA ring
    must not be handed back to the producer before the fragment loads of its k-tile have completed; when it was
    (mbarrier.arrive scheduled above the DMMAs consuming the loads), the 4-consumer-warp configurations returned a
    handful of wrong entries in some launches -- every launch is checked, entry by entry."""
    import torch
    from hermipy_b200.plan import dgemm
    g = torch.Generator(device="cuda").manual_seed(5)
    for (M, N, K, batch) in [(128, 256, 128, 512), (128, 4096, 128, 6)]:
        A = torch.randn((batch, M, K), generator=g, dtype=torch.float64, device="cuda")
        B = torch.randn((batch, K, N), generator=g, dtype=torch.float64, device="cuda")
        want = torch.matmul(A, B)
        os.environ["HB_GEMM_CFG"] = str(cfg)
        try:
            for rep in range(6):
                C = torch.full((batch, M, N), float("nan"), dtype=torch.float64, device="cuda")
                dgemm(M, N, K, A, K, B, N, C, N, batch=batch, strideA=M * K, strideB=K * N, strideC=M * N)
                wrong = int(((C - want).abs() > 1e-10).sum() + (~torch.isfinite(C)).sum())
                assert wrong == 0, (cfg, M, N, K, batch, rep, wrong)
        finally:
            os.environ["HB_GEMM_CFG"] = "-1"


@pytest.mark.parametrize("cfg", [-1, 0, 1, 2, 4, 7, 8])
def test_two_level_batch_and_row_block_scatter(cfg):
    """hb_dev_dgemm_ex: batch * batch_lo GEMMs in one launch (per-level strides, a table shared across b_hi), plain
    store and the row-block scatter with row_offset_lo, against torch.matmul."""
    import torch
    from hermipy_b200.dist import DeviceOps, Gemm
    ops = DeviceOps()
    g = torch.Generator(device="cuda").manual_seed(9)
    os.environ["HB_GEMM_CFG"] = str(cfg)
    try:
        for (M, N, K, nhi, nlo) in [(128, 256, 128, 40, 2), (101, 202, 50, 3, 2), (64, 130, 33, 5, 3)]:
            Kp, Np = (K + 1) & ~1, (N + 1) & ~1
            A = torch.randn((nlo, M, Kp), generator=g, dtype=torch.float64, device="cuda")
            B = torch.randn((nhi, nlo, K, Np), generator=g, dtype=torch.float64, device="cuda")
            want = torch.einsum("lmk,hlkn->hlmn", A[:, :, :K], B[..., :N])
            C = torch.full((nhi, nlo, M, Np), float("nan"), dtype=torch.float64, device="cuda")
            ops.gemm(Gemm(M, N, K, A, Kp, B, Np, C, Np, batch=nhi, batch_lo=nlo, strideA_lo=M * Kp, strideB=nlo * K * Np,
                          strideB_lo=K * Np, strideC=nlo * M * Np, strideC_lo=M * Np))
            assert (C[..., :N] - want).abs().max().item() < 1e-11
            # scatter: destination row = l * M + m, dealt over nlo blocks of M rows: D[l][h][m][n]
            D = torch.full((nlo, nhi, M, Np), float("nan"), dtype=torch.float64, device="cuda")
            ops.gemm(Gemm(M, N, K, A, Kp, B, Np, ldc=Np, batch=nhi, batch_lo=nlo, strideA_lo=M * Kp, strideB=nlo * K * Np,
                          strideB_lo=K * Np, strideC=M * Np, blocks=[(D, l * nhi * M * Np) for l in range(nlo)],
                          block_rows=M, row_offset_lo=M))
            assert (D.permute(1, 0, 2, 3)[..., :N] - want).abs().max().item() < 1e-11
    finally:
        os.environ["HB_GEMM_CFG"] = "-1"


@pytest.mark.parametrize("cfg", [1, 2, 6, 7])
def test_bulk_store_epilogue_equals_register_epilogue(cfg):
    """The row-block scatter leaves through shared memory and bulk asynchronous copies (EPI_BULK) when the layout
    allows 16-byte runs: same bits as the register epilogue (HB_GEMM_BULK=0), persistent CTAs with several tiles
    each (the staging buffer is reused), edge tiles in M and N, and nothing written beyond column N."""
    import torch
    from hermipy_b200.dist import DeviceOps, Gemm
    ops = DeviceOps()
    g = torch.Generator(device="cuda").manual_seed(11)
    os.environ["HB_GEMM_CFG"] = str(cfg)
    try:
        for (M, N, K, nhi, nlo, ld) in [(128, 256, 128, 64, 2, 256), (200, 1000, 48, 7, 2, 1010), (72, 44, 130, 9, 1, 48)]:
            Kp = (K + 1) & ~1
            A = torch.randn((nlo, M, Kp), generator=g, dtype=torch.float64, device="cuda")
            B = torch.randn((nhi, nlo, K, N), generator=g, dtype=torch.float64, device="cuda")
            res = []
            for bulk in ("1", "0"):
                os.environ["HB_GEMM_BULK"] = bulk
                D = torch.full((nlo, nhi, M, ld), float("nan"), dtype=torch.float64, device="cuda")
                ops.gemm(Gemm(M, N, K, A, Kp, B, N, ldc=ld, batch=nhi, batch_lo=nlo, strideA_lo=M * Kp, strideB=nlo * K * N,
                              strideB_lo=K * N, strideC=M * ld, blocks=[(D, l * nhi * M * ld) for l in range(nlo)],
                              block_rows=M, row_offset_lo=M))
                res.append(D)
            want = torch.einsum("lmk,hlkn->lhmn", A[:, :, :K], B)
            assert torch.equal(res[0][..., :N], res[1][..., :N])
            assert (res[0][..., :N] - want).abs().max().item() < 1e-10
            assert torch.isnan(res[0][..., N:]).all()
    finally:
        os.environ["HB_GEMM_CFG"] = "-1"
        os.environ.pop("HB_GEMM_BULK", None)
