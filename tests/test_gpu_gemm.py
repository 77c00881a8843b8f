"""tcgen05 GEMM core (through the C ABI) against a plain fp32 torch reference of the same contraction."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_major, b_major):
    A = a.float() if a_major == 0 else a.float().t()
    B = b.float() if b_major == 0 else b.float().t()
    return A @ B.t()


@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (384, 768, 768), (200, 136, 72), (128, 2304, 768)])
def test_gemm_majors_bf16_out(a_major, b_major, M, N, K):
    from contrastors_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    a = torch.randn((M, K) if a_major == 0 else (K, M), device=dev).to(torch.bfloat16)
    b = torch.randn((N, K) if b_major == 0 else (K, N), device=dev).to(torch.bfloat16)
    c = ops.gemm(a, b, a_major=a_major, b_major=b_major)
    ref = _ref(a, b, a_major, b_major)
    err = (c.float() - ref).abs().max().item()
    # bf16 output rounding: 2^-9 relative on values up to ~4*sqrt(K)
    assert err <= 2.0 ** -8 * ref.abs().max().item() + 1e-3, err


@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(512, 256, 64), (1024, 768, 768), (4096, 2304, 320), (2048, 768, 6144)])
def test_gemm_quad_cluster_multicast_equals_pair_mode(a_major, b_major, M, N, K):
    """M % 512 == 0 selects the 4-CTA cluster (two CTA pairs, B halves fetched once per cluster by TMA multicast): the same
    MMAs on the same operands as pair mode, so the outputs must be BITWISE equal, and both match the fp32 reference."""
    from contrastors_b200 import ops
    torch.manual_seed(7)
    a = torch.randn((M, K) if a_major == 0 else (K, M), device="cuda").to(torch.bfloat16)
    b = torch.randn((N, K) if b_major == 0 else (K, N), device="cuda").to(torch.bfloat16)
    try:
        ops.gemm_select_cluster(2)
        pair = ops.gemm(a, b, a_major=a_major, b_major=b_major)
        pair32 = ops.gemm(a, b, a_major=a_major, b_major=b_major, out_dtype=torch.float32)
        ops.gemm_select_cluster(4)
        quad = ops.gemm(a, b, a_major=a_major, b_major=b_major)
        quad32 = ops.gemm(a, b, a_major=a_major, b_major=b_major, out_dtype=torch.float32)
    finally:
        ops.gemm_select_cluster(0)
    ref = _ref(a, b, a_major, b_major)
    assert (quad.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() + 1e-3
    assert torch.equal(pair, quad)
    # fp32: split-K partials are reduce-added in arrival order, so only near-equality
    assert (quad32 - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() * (K ** 0.5) + 1e-4
    assert (pair32 - quad32).abs().max().item() <= 1e-5 * ref.abs().max().item() * (K ** 0.5) + 1e-4


def test_gemm_swiglu_quad_cluster():
    import torch.nn.functional as F
    from contrastors_b200 import ops
    torch.manual_seed(8)
    M, I, K = 1024, 3072, 768
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(2 * I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    try:
        ops.gemm_select_cluster(2)
        act_p, yg_p = ops.gemm_swiglu(x, w1, keep_preact=True)
        ops.gemm_select_cluster(4)
        act_q, yg_q = ops.gemm_swiglu(x, w1, keep_preact=True)
    finally:
        ops.gemm_select_cluster(0)
    ref_yg = x.float() @ w1.float().t()
    assert (yg_q.float() - ref_yg).abs().max().item() <= 2.0 ** -8 * ref_yg.abs().max().item() + 1e-3
    assert torch.equal(act_p, act_q) and torch.equal(yg_p, yg_q)


@pytest.mark.parametrize("a_major,b_major,M,N,K", [(0, 0, 256, 256, 4096), (1, 1, 768, 768, 8192), (0, 1, 2048, 768, 16384),
                                                    (1, 1, 16384, 768, 2048), (1, 0, 136, 72, 200)])
def test_gemm_fp32_out_splitk_and_accumulate(a_major, b_major, M, N, K):
    from contrastors_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    a = torch.randn((M, K) if a_major == 0 else (K, M), device=dev).to(torch.bfloat16)
    b = torch.randn((N, K) if b_major == 0 else (K, N), device=dev).to(torch.bfloat16)
    ref = _ref(a, b, a_major, b_major)
    c = ops.gemm(a, b, a_major=a_major, b_major=b_major, out_dtype=torch.float32, alpha=0.5)
    tol = 1e-5 * ref.abs().max().item() * (K ** 0.5) + 1e-4
    assert (c - 0.5 * ref).abs().max().item() <= tol
    base = torch.randn(M, N, device=dev)
    out = base.clone()
    ops.gemm(a, b, a_major=a_major, b_major=b_major, out=out, accumulate=True)
    assert (out - (base + ref)).abs().max().item() <= tol


def test_gemm_strided_operands():
    from contrastors_b200 import ops
    torch.manual_seed(2)
    big_a = torch.randn(300, 1024, device="cuda").to(torch.bfloat16)
    big_b = torch.randn(520, 1024, device="cuda").to(torch.bfloat16)
    a, b = big_a[:, :256], big_b[:, :256]  # K prefix of a wider row stride (the Matryoshka access pattern)
    c = ops.gemm(a, b, out_dtype=torch.float32)
    ref = a.float() @ b.float().t()
    assert (c - ref).abs().max().item() <= 1e-3


@pytest.mark.parametrize("M,I,K", [(256, 128, 64), (1000, 3072, 768), (130, 256, 128)])
def test_gemm_swiglu_epilogue(M, I, K):
    """fc11/fc12 GEMM with the SwiGLU fused into the epilogue (layers/mlp.py:68-75) vs fp32 torch."""
    import torch.nn.functional as F
    from contrastors_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(2 * I, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    act, yg = ops.gemm_swiglu(x, w1, keep_preact=True)
    ref_yg = x.float() @ w1.float().t()
    ref = ref_yg[:, :I] * F.silu(ref_yg[:, I:])
    assert (yg.float() - ref_yg).abs().max().item() <= 2.0 ** -8 * ref_yg.abs().max().item() + 1e-3
    assert (act.float() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item() + 1e-3
    act2, none = ops.gemm_swiglu(x, w1, keep_preact=False)
    assert none is None and torch.equal(act, act2)


@pytest.mark.parametrize("M,I,K", [(256, 256, 64), (1024, 3072, 768), (1000, 512, 192)])
def test_gemm_swiglu_bwd_epilogue(M, I, K):
    """fc2 dgrad with the SwiGLU backward in its epilogue vs fp32 torch autograd of y * silu(gate), and vs the two-kernel form
    (dgrad GEMM rounded to bf16, then cx_swiglu_bwd) it replaces."""
    import torch.nn.functional as F
    from contrastors_b200 import ops
    torch.manual_seed(9)
    dout = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w2 = (torch.randn(K, I, device="cuda") / K ** 0.5).to(torch.bfloat16)
    yg = torch.randn(M, 2 * I, device="cuda").to(torch.bfloat16)
    got = ops.gemm_swiglu_bwd(dout, w2, yg)
    ygf = yg.float().requires_grad_()
    act = ygf[:, :I] * F.silu(ygf[:, I:])
    act.backward(dout.float() @ w2.float())
    ref = ygf.grad
    assert (got.float() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item() + 1e-3
    two = ops.swiglu_bwd(ops.gemm(dout, w2, b_major=1), yg)
    assert (got.float() - two.float()).abs().max().item() <= 2.0 ** -6 * ref.abs().max().item() + 1e-3


def test_gemm_qkv_rope_epilogue_matches_gemm_then_rope():
    """RoPE fused into the QKV GEMM epilogue == plain GEMM followed by the standalone rotary kernel (bitwise up to the
    single bf16 rounding the fusion removes), and both match an fp32 torch reference."""
    from contrastors_b200 import ops
    torch.manual_seed(4)
    H, Dh, d = 3, 64, 192
    lens = [200, 56]
    T = sum(lens)
    cu = torch.tensor([0, 200, 256], dtype=torch.int32, device="cuda")
    pos = ops.token_positions(cu, T)
    inv = 1.0 / (1000.0 ** (torch.arange(0, Dh, 2, dtype=torch.float32) / Dh))
    fr = torch.outer(torch.arange(256, dtype=torch.float32), inv)
    cos_t, sin_t = torch.cos(fr).cuda(), torch.sin(fr).cuda()
    x = torch.randn(T, d, device="cuda").to(torch.bfloat16)
    w = (torch.randn(3 * d, d, device="cuda") / d ** 0.5).to(torch.bfloat16)
    fused = ops.gemm_qkv_rope(x, w, pos, inv.cuda(), 2 * d)
    ref = (x.float() @ w.float().t()).view(T, 3, H, Dh)
    c = cos_t[pos.long()][:, None, :]
    s = sin_t[pos.long()][:, None, :]
    for slot in (0, 1):
        x1, x2 = ref[:, slot, :, :32].clone(), ref[:, slot, :, 32:].clone()
        ref[:, slot, :, :32] = x1 * c - x2 * s
        ref[:, slot, :, 32:] = x2 * c + x1 * s
    assert (fused.float().view(T, 3, H, Dh) - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()
    two_step = ops.rope_inplace(ops.gemm(x, w), pos, cos_t, sin_t, H, Dh)
    assert (fused.float() - two_step.float()).abs().max().item() <= 2.0 ** -6 * ref.abs().max().item()
