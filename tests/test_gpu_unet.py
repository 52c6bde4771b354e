"""GPU parity of the CQT and the whole denoiser forward against the CPU oracle (oracle/), through the C ABI.

Tolerance: rel-L2 <= 1e-3 per denoiser call; the exact-fp32 path (conv_mode 0) is held to
1e-4, two orders above its observed error, so that a broken kernel cannot hide behind the published bar.
"""
import pytest
import torch

from util import rel_l2, seeded, make_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(aid, cuda):
    cfg = aid.small_test(16384)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    net.to(cuda)
    return cfg, sd, net, make_oracle(cfg, sd)


@pytest.mark.parametrize("L", [16384, 65536])
def test_cqt_fwd_bwd_hpf(aid, cuda, L):
    cfg = aid.small_test(L)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    import cqt_oracle
    orc = cqt_oracle.CQT_nsgt(cfg.num_octs, cfg.bins_per_oct, "oct", ("kaiser", cfg.beta), fs=cfg.sample_rate, audio_len=L)
    x = seeded((3, 1, L), 11)
    X_ref = orc.fwd(x)
    X = net.CQTransform.fwd(x.to(cuda))
    assert [tuple(a.shape) for a in X] == [tuple(a.shape) for a in X_ref]
    for a, b in zip(X, X_ref):
        assert rel_l2(torch.view_as_real(a), torch.view_as_real(b)) < 2e-6
    # synthesis from independent random coefficients (not only from an analysis)
    coefs = [torch.complex(seeded(tuple(c.shape), 20 + i), seeded(tuple(c.shape), 40 + i)) for i, c in enumerate(X_ref)]
    y_ref = orc.bwd(coefs)
    y = net.CQTransform.bwd([c.to(cuda) for c in coefs])
    assert y.shape == y_ref.shape
    assert rel_l2(y, y_ref) < 2e-6
    h_ref = orc.apply_hpf_DC(x[:, 0])
    hh = net.CQTransform.apply_hpf_DC(x[:, 0].to(cuda))
    assert rel_l2(hh, h_ref) < 2e-6
    # shorter input is zero padded (sampler.py:63 is called on whatever length the sampler carries)
    hs = net.CQTransform.apply_hpf_DC(x[:, 0, : L - 100].to(cuda))
    assert rel_l2(hs, orc.apply_hpf_DC(x[:, 0, : L - 100])) < 2e-6


def test_cqt_roundtrip_property(aid, cuda):
    """bwd(fwd(x)) == apply_hpf_DC(x) for a signal with no energy in the top 2 % of the band (frame identity)."""
    L = 65536
    cfg = aid.small_test(L)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    x = seeded((2, L), 5)
    X = torch.fft.rfft(x)
    X[:, int(0.97 * (L // 2)):] = 0
    x = torch.fft.irfft(X, n=L).to(cuda)
    y = net.CQTransform.bwd(net.CQTransform.fwd(x.unsqueeze(1)))[:, 0]
    assert rel_l2(y, net.CQTransform.apply_hpf_DC(x)) < 1e-5


@pytest.mark.parametrize("cnoise", [0.0613, -0.75, -2.3])
def test_forward_small_vs_oracle(small, cuda, cnoise):
    cfg, sd, net, orc = small
    x = seeded((2, cfg.audio_len), 0)
    cn = torch.tensor([[cnoise]])
    ref = orc(x, cn)
    out = net(x.to(cuda), cn.to(cuda))
    assert out.shape == x.shape and out.dtype == torch.float32
    assert rel_l2(out, ref) < 1e-4


def test_forward_per_sample_sigma_and_batch_independence(small, cuda):
    cfg, sd, net, orc = small
    x = seeded((3, cfg.audio_len), 3, 0.3)
    cn = torch.tensor([[0.05], [-0.8], [-1.7]])
    ref = orc(x, cn)
    out = net(x.to(cuda), cn.to(cuda))
    assert rel_l2(out, ref) < 1e-4
    # clips are independent: row 1 alone gives the same answer as row 1 inside the batch
    solo = net(x[1:2].to(cuda), cn[1:2].to(cuda))
    assert rel_l2(solo, out[1:2]) < 1e-5


def test_forward_fused_preconditioning(small, cuda):
    """out = c_skip*x + c_out*net(c_in*x, c_noise)  (edm.py:133-148) in one call."""
    cfg, sd, net, orc = small
    import unet_oracle
    edm = unet_oracle.EDMOracle()
    x = seeded((2, cfg.audio_len), 9, 0.5)
    sigma = torch.tensor([0.4])
    ref = edm.denoiser(x, orc, sigma)
    s = sigma.reshape(1, 1)
    sdv = edm.sigma_data
    cskip, cout, cin = sdv ** 2 / (s ** 2 + sdv ** 2), s * sdv * (sdv ** 2 + s ** 2) ** -0.5, (sdv ** 2 + s ** 2) ** -0.5
    out = net.denoise_fused(x.to(cuda), 0.25 * torch.log(s), float(cin), float(cout), float(cskip))
    assert rel_l2(out, ref) < 1e-4


def test_forward_requires_grad_raises(small, cuda):
    cfg, sd, net, orc = small
    x = seeded((1, cfg.audio_len), 1).to(cuda).requires_grad_()
    with pytest.raises(RuntimeError, match="forward-only"):
        net(x, torch.tensor([[0.0]], device=cuda))
    with torch.no_grad():
        net(x, torch.tensor([[0.0]], device=cuda))


def test_forward_paper_network_config1(aid, cuda):
    """BASELINE config 1: the 186 M-parameter network, 1 x 65536, through EDM.denoiser at sigma = 1."""
    import unet_oracle
    cfg = aid.paper_22k(65536)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    orc = make_oracle(cfg, sd)
    edm = unet_oracle.EDMOracle()
    x = seeded((1, 65536), 0)
    sigma = torch.tensor([1.0])
    ref = edm.denoiser(x, orc, sigma)
    e = aid.EDM(aid.AttrDict.wrap({"diff_params": dict(sigma_min=1e-4, sigma_max=1.0, ro=13, sigma_data=0.063, Schurn=10,
                                                        Stmin=0, Stmax=50, Snoise=1.0)}))
    out = e.denoiser(x.to(cuda), net, sigma.to(cuda))
    assert rel_l2(out, ref) < 1e-4
