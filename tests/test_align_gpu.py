"""GPU tests of the fused kernels around the alignment search (csrc/align.cu): each one
against the oracle's restatement of the reference's expression (oracle/glow_oracle.py: log_prior, mle_loss;
Modules.py:107-122, 1020-1029 -- pure torch, so it is evaluated in fp32 on the same device and inputs).  Tolerances: fp32 sums in a different order -> 1e-5 of the largest magnitude; integer outputs exact."""
import math

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

T_LENS = [61, 12, 150, 97, 202, 33]
M_LENS = [400, 104, 946, 610, 1000, 250]


def _inputs(seed=0, c=80):
    g = torch.Generator().manual_seed(seed)
    b, tx, ty = len(T_LENS), max(T_LENS), max(M_LENS)
    dev = torch.device("cuda:0")
    mean = (torch.randn(b, c, tx, generator=g) * 0.7).to(dev)
    log_std = (torch.randn(b, c, tx, generator=g) * 0.3 - 0.2).to(dev)
    z = torch.randn(b, c, ty, generator=g).to(dev)
    t_len = torch.tensor(T_LENS, dtype=torch.int32, device=dev)
    m_len = torch.tensor(M_LENS, dtype=torch.int32, device=dev)
    tmask = (torch.arange(tx, device=dev)[None] < t_len[:, None]).float().unsqueeze(1)
    mmask = (torch.arange(ty, device=dev)[None] < m_len[:, None]).float().unsqueeze(1)
    return mean * tmask, log_std * tmask, z * mmask, t_len, m_len, tmask, mmask


def _ref_log_p(z, mean, log_std):
    """The oracle's restatement of Modules.py:107-116 (oracle/glow_oracle.py:log_prior, pure torch), evaluated on the
    device the kernel ran on."""
    from oracle import glow_oracle
    return glow_oracle.log_prior(z, mean, log_std)


def test_log_p_matches_reference_expression_on_the_valid_corner():
    from glow_tts_b200 import align
    mean, log_std, z, t_len, m_len, tmask, mmask = _inputs()
    torch.backends.cuda.matmul.allow_tf32 = False
    want = _ref_log_p(z, mean, log_std)
    got = align.log_p(z, mean, log_std, t_len, m_len)
    corner = (tmask.transpose(1, 2) * mmask) > 0
    err = (got - want)[corner].abs().max().item()
    assert err < 1e-5 * want[corner].abs().max().item(), err


def test_search_on_fused_log_p_finds_the_same_paths_and_reports_tokens_and_durations():
    from glow_tts_b200 import align
    from glow_tts_b200.monotonic_align import maximum_path, maximum_path_align
    mean, log_std, z, t_len, m_len, tmask, mmask = _inputs(1)
    want_lp = _ref_log_p(z, mean, log_std)
    lp = align.log_p(z, mean, log_std, t_len, m_len)
    path_ref = maximum_path(want_lp, None, t_len, m_len)
    path, tok, dur = maximum_path_align(lp, t_len, m_len)
    # continuous random values: no near-ties, so the differently rounded log_P must give the same alignment
    assert torch.equal(path, path_ref)
    assert torch.equal(dur, path.sum(-1).to(torch.int32))
    ty = path.shape[2]
    valid = torch.arange(ty, device=path.device)[None] < m_len[:, None]
    assert torch.equal(tok[valid], path.argmax(dim=1).to(torch.int32)[valid])
    assert int(tok[~valid].abs().sum()) == 0
    assert torch.equal(dur.sum(-1), m_len)


def test_expand_by_path_forward_and_backward_match_the_dense_products():
    from glow_tts_b200 import align
    from glow_tts_b200.monotonic_align import maximum_path_align
    mean, log_std, z, t_len, m_len, tmask, mmask = _inputs(2)
    path, tok, dur = maximum_path_align(_ref_log_p(z, mean, log_std), t_len, m_len)
    m1, s1 = mean.clone().requires_grad_(True), log_std.clone().requires_grad_(True)
    m2, s2 = mean.clone().requires_grad_(True), log_std.clone().requires_grad_(True)
    want_m, want_s = m1 @ path, s1 @ path                                    # Modules.py:120-121
    want_t = torch.log(path.sum(dim=-1).unsqueeze(1) + 1e-7) * tmask         # Modules.py:122
    got_m, got_s, got_t = align.expand_by_path(m2, s2, tok, dur, t_len, m_len, z.shape[2])
    assert torch.equal(got_m, want_m) and torch.equal(got_s, want_s)          # a one-term sum: bit-identical
    assert (got_t - want_t).abs().max().item() < 1e-6
    gm, gs = torch.randn_like(want_m), torch.randn_like(want_s)
    torch.autograd.backward([want_m, want_s], [gm, gs])
    torch.autograd.backward([got_m, got_s], [gm, gs])
    for a, b in ((m2.grad, m1.grad), (s2.grad, s1.grad)):
        assert (a - b).abs().max().item() < 1e-5 * b.abs().max().item()


def test_mle_loss_forward_and_backward_match_the_reference_expression():
    from glow_tts_b200 import align
    mean, log_std, z, t_len, m_len, tmask, mmask = _inputs(3)
    b, c, ty = z.shape
    mel_mean = (torch.randn(b, c, ty, device=z.device) * 0.5) * mmask
    mel_std = (torch.randn(b, c, ty, device=z.device) * 0.3) * mmask
    log_dets = torch.randn(b, device=z.device) * 50
    lengths = m_len.long()

    def ref(z, m, s, ld):                                                     # oracle/glow_oracle.py:mle_loss = Modules.py:1020-1029
        from types import SimpleNamespace
        from oracle import glow_oracle
        return glow_oracle.mle_loss(z, m, s, ld, lengths, SimpleNamespace(num_squeeze=2, mel_dim=80))

    a = [t.clone().requires_grad_(True) for t in (z, mel_mean, mel_std, log_dets)]
    g = [t.clone().requires_grad_(True) for t in (z, mel_mean, mel_std, log_dets)]
    want = ref(*a)
    got = align.mle_loss(g[0], g[1], g[2], g[3], lengths, 2, 80)
    assert abs(float(got) - float(want)) < 1e-5 * abs(float(want))
    (want * 1.7).backward()
    (got * 1.7).backward()
    for x, y in zip(g, a):
        assert (x.grad - y.grad).abs().max().item() < 1e-5 * y.grad.abs().max().item()
