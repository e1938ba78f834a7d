"""fp32 mode (BASELINE.json north_star: "relative L2 at most 1e-5 per step in fp32 mode"; SURVEY 8d config C1: CIFAR-10
DDPM U-Net T=10 + energy eval, batch 8, fp32).  The CUDA-core fp32 path (csrc/kernels_f32.cu, csrc/engine_f32.cu) against the
reference's own fp32 outputs (tests/golden, generated from /root/reference on the CPU) and against the CPU oracle at B=8.

Tolerances asserted here:
  * per step, teacher-forced on the reference's x_t (no accumulation):  rel-L2(x_{t+1}) <= 1e-5   (the north_star bound)
  * the raw U-Net output eps on the same input:                         rel-L2 <= 1e-5
  * free-running T=10 rollout (errors compound over 10 U-Net calls):    rel-L2 <= 1e-5 on every x_t
(measured on B200: eps 2.2e-6, per step <= 8.2e-7, free-running <= 8.2e-7, energy 1.3e-7)
"""
import pytest
import torch

from common import build_ddpm, golden, rel_l2

pytestmark = pytest.mark.gpu
TOL_STEP = 1e-5
TOL_ROLLOUT = 1e-5


@pytest.fixture(autouse=True)
def _inference_mode():
    with torch.no_grad():  # the fp32 mode is inference-only
        yield


def test_fp32_value_net_refuses_autograd(ddpm10_fp32):
    net, sampler, value, sd, vsd = ddpm10_fp32
    with torch.enable_grad(), pytest.raises(RuntimeError, match="inference-only"):
        value(torch.randn(2, 3, 32, 32, device="cuda"), 0)


@pytest.fixture(scope="module")
def ddpm10_fp32():
    net, sampler, value, sd, vsd = build_ddpm(10)
    net.set_precision("fp32")
    value.net.set_precision("fp32")
    assert net.precision == "fp32" and value.net.precision == "fp32"
    return net, sampler, value, sd, vsd


def test_fp32_unet_forward_matches_reference(ddpm10_fp32):
    net, sampler, value, sd, vsd = ddpm10_fp32
    g = golden("ddpm_T10_B2.npz")
    x0 = torch.from_numpy(g["l_sample"][0]).cuda()
    t = torch.full((2,), float(g["continuous_steps"][0]), device="cuda")
    eps = net(x0, t)
    torch.cuda.synchronize()
    err = rel_l2(eps, torch.from_numpy(g["eps_first"]))
    print("fp32 eps rel-L2 %.3e" % err)
    assert err < TOL_STEP


def test_fp32_teacher_forced_steps(ddpm10_fp32):
    from oracle import synth

    net, sampler, value, sd, vsd = ddpm10_fp32
    g = golden("ddpm_T10_B2.npz")
    noise = synth.synth_noise(10, 2, (3, 32, 32))
    errs = []
    for i in range(10):
        x = torch.from_numpy(g["l_sample"][i]).cuda()
        d = sampler.sample_step(x, i, noise=noise[i + 1])
        errs.append(rel_l2(d["sample"], torch.from_numpy(g["l_sample"][i + 1])))
    print("fp32 per-step rel-L2:", ["%.2e" % e for e in errs])
    assert max(errs) < TOL_STEP


def test_fp32_rollout_and_energy(ddpm10_fp32):
    from oracle import synth

    net, sampler, value, sd, vsd = ddpm10_fp32
    g = golden("ddpm_T10_B2.npz")
    noise = torch.stack(synth.synth_noise(10, 2, (3, 32, 32)))
    d = sampler.sample(2, device="cuda", noise=noise)
    errs = [rel_l2(d["l_sample"][i], torch.from_numpy(g["l_sample"][i])) for i in range(11)]
    print("fp32 cumulative rel-L2:", ["%.2e" % e for e in errs])
    assert errs[0] == 0.0 and max(errs) < TOL_ROLLOUT
    lp = torch.stack(d["logp"]).cpu()
    assert torch.allclose(lp, torch.from_numpy(g["logp"]), atol=1e-5, rtol=1e-5)
    # energy of the reference's own final sample (teacher-forced) and of ours
    e_ref_x = value(torch.from_numpy(g["l_sample"][10]).cuda(), 10)
    err_e = rel_l2(e_ref_x, torch.from_numpy(g["energy"]))
    print("fp32 energy rel-L2 %.3e" % err_e)
    assert err_e < TOL_STEP
    assert rel_l2(value(d["sample"], 10), torch.from_numpy(g["energy"])) < TOL_ROLLOUT


def test_fp32_config_c1_batch8_vs_oracle(ddpm10_fp32):
    """BASELINE.json configs[0]: T=10, batch 8, fp32 - fresh seeded inputs, CPU oracle vs the GPU fp32 mode, per step."""
    from oracle import nets, samplers, synth

    net, sampler, value, sd, vsd = ddpm10_fp32
    B = 8
    noise = synth.synth_noise(10, B, (3, 32, 32), seed=11)
    sched = samplers.var_schedule(10)
    with torch.no_grad():
        ref = samplers.var_rollout(lambda x, t: nets.ddpm_unet_forward(sd, x, t), sched, sd["log_betas"], noise)
        eref = nets.value_forward(vsd, ref["sample"])
    errs = []
    for i in range(10):
        d = sampler.sample_step(ref["l_sample"][i].cuda(), i, noise=noise[i + 1])
        errs.append(rel_l2(d["sample"], ref["l_sample"][i + 1]))
    print("C1 B=8 per-step rel-L2:", ["%.2e" % e for e in errs])
    assert max(errs) < TOL_STEP
    assert rel_l2(value(ref["sample"].cuda(), 10), eref) < TOL_STEP
    full = sampler.sample(B, device="cuda", noise=torch.stack(noise))
    assert max(rel_l2(full["l_sample"][i], ref["l_sample"][i]) for i in range(11)) < TOL_ROLLOUT


def test_fp32_batch_invariance_and_bf16_agreement(ddpm10_fp32):
    net, sampler, value, sd, vsd = ddpm10_fp32
    x = torch.randn(5, 3, 32, 32, device="cuda")
    t = torch.full((5,), 77.0, device="cuda")
    a = net(x, t)
    b = net(x[:2].contiguous(), t[:2])
    assert torch.equal(a[:2], b)  # fixed reduction orders: bitwise independent of the batch size
    net.set_precision("bf16")
    try:
        c = net(x, t)
    finally:
        net.set_precision("fp32")
    err = rel_l2(c, a)
    print("bf16 vs fp32 mode eps rel-L2 %.3e" % err)
    assert 1e-6 < err < 2e-2


def test_fp32_mode_rejects_adm():
    from common import EDM_SMALL_CFG, build_edm

    unet, sampler, sd = build_edm(EDM_SMALL_CFG, 2)
    unet.set_precision("fp32")
    x = torch.randn(2, 3, 32, 32, device="cuda")
    with pytest.raises(RuntimeError, match="fp32 mode"):
        unet(x, torch.full((2,), 1.0, device="cuda"), torch.zeros(2, dtype=torch.long, device="cuda"))
