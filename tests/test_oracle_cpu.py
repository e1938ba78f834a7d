"""CPU suite, part 1: the oracle against the reference's golden vectors, and the product's host-side schedule logic
against the oracle.  No GPU, no /root/reference (fixtures under tests/golden were generated from it by
oracle/gen_golden.py)."""
import math

import numpy as np
import torch

from common import golden, rel_l2
from oracle import nets, samplers, synth

# The only known-answer constant in the reference tree: models/DxMI/trainer.py:147-149.
ETA_T10 = [1.00000e-04, 1.10250e-02, 4.00000e-02, 8.70250e-02, 1.52100e-01, 2.35225e-01, 3.36400e-01, 4.55625e-01,
           5.92900e-01, 7.48225e-01]


def test_user_defined_eta_known_answer():
    eta = samplers.var_eta(10)
    assert np.allclose(eta, ETA_T10, rtol=2e-5, atol=0)
    from diffusion_by_maxentirl_b200.schedule import quadratic_eta

    assert np.array_equal(quadratic_eta(10), eta)


def _check_var_tables(T, name):
    g = golden(name)
    sched = samplers.var_schedule(T)
    for k in ("continuous_steps", "Gamma_bar", "x_prev_multiplier", "theta_multiplier", "std"):
        assert np.array_equal(sched[k].numpy(), g[k]), k
    assert np.array_equal(sched["user_defined_eta"], g["user_defined_eta"])
    assert np.array_equal(sched["log_betas_init"].numpy(), g["log_betas"])
    # product host logic (schedule.py) == oracle == reference buffers, bit for bit
    from diffusion_by_maxentirl_b200.schedule import VarSchedule

    s = VarSchedule(T)
    for k in ("continuous_steps", "Gamma_bar", "x_prev_multiplier", "theta_multiplier", "std"):
        assert np.array_equal(getattr(s, k).numpy(), g[k]), k
    assert abs(float(s.continuous_steps[-1])) < 0.1  # var_sampler.py:256
    return g, sched


def test_var_schedule_matches_reference_T10():
    _check_var_tables(10, "ddpm_T10_B2.npz")


def test_var_schedule_matches_reference_T4():
    _check_var_tables(4, "ddpm_T4_B2.npz")


def _ddpm_weights():
    import json
    import os

    from common import GOLD

    shapes = json.load(open(os.path.join(GOLD, "ddpm_shapes.json")))
    sd = synth.synth_state_dict({k: tuple(v) for k, v in shapes["net"].items()})
    vsd = synth.synth_state_dict({k: tuple(v) for k, v in shapes["value"].items()}, seed=1)
    return sd, vsd


def test_oracle_rollout_reproduces_reference_T4():
    """The restated U-Net + sampler + value net give the reference's own outputs (fp32 round-off)."""
    g, sched = _check_var_tables(4, "ddpm_T4_B2.npz")
    sd, vsd = _ddpm_weights()
    noise = synth.synth_noise(4, 2, (3, 32, 32))
    with torch.no_grad():
        d = samplers.var_rollout(lambda x, t: nets.ddpm_unet_forward(sd, x, t), sched, sched["log_betas_init"], noise)
        e = nets.value_forward(vsd, d["sample"])
    for i in range(5):
        assert rel_l2(d["l_sample"][i], torch.from_numpy(g["l_sample"][i])) < 2e-6
    assert rel_l2(torch.stack(d["logp"]), torch.from_numpy(g["logp"])) < 2e-5
    assert rel_l2(d["mean"][-1], torch.from_numpy(g["mean_last"])) < 2e-6
    assert rel_l2(d["control"][0], torch.from_numpy(g["control_first"])) < 2e-6
    assert rel_l2(d["eps"][0], torch.from_numpy(g["eps_first"])) < 2e-6
    assert rel_l2(e, torch.from_numpy(g["energy"])) < 2e-6
    # sample() == looped sample_step() (two reference code paths for the same math)
    with torch.no_grad():
        x = noise[0]
        for i in range(4):
            o = samplers.var_sample_step(lambda x, t: nets.ddpm_unet_forward(sd, x, t), sched, sched["log_betas_init"],
                                         x, torch.full((2,), i, dtype=torch.long), noise[i + 1])
            x = o["sample"]
            assert o["sigma"].shape == (2, 1, 1, 1) and o["logp"].shape == (2,)
    assert rel_l2(x, d["sample"]) < 1e-6


def test_var_transition_edge_cases():
    """fix_last pins the last sigma to std[-1] = 1e-3; trainable_beta=False uses the schedule sigma; the mean uses the
    schedule sigma even when log_betas was trained away from it."""
    sched = samplers.var_schedule(4)
    lb = sched["log_betas_init"] + 0.3
    s_fix = samplers.var_sigmas(lb, sched["std"], "fix_last")
    assert torch.allclose(s_fix[:-1], torch.exp(lb[:-1])) and abs(float(s_fix[-1]) - 1e-3) < 1e-9
    assert torch.equal(samplers.var_sigmas(lb, sched["std"], False), sched["std"])
    assert torch.allclose(samplers.var_sigmas(lb, sched["std"], True), torch.exp(lb))
    noise = synth.synth_noise(4, 3, (3, 8, 8))
    zero_net = lambda x, t: torch.zeros_like(x)
    d0 = samplers.var_rollout(zero_net, sched, sched["log_betas_init"], noise)
    d1 = samplers.var_rollout(zero_net, sched, lb, noise)
    assert torch.equal(d0["mean"][0], d1["mean"][0])  # same mean, different noise scale
    assert not torch.equal(d0["l_sample"][1], d1["l_sample"][1])
    # log-prob of the drawn sample only depends on z: mean(-z^2/2) - ln sigma - 0.5 ln 2 pi
    z = noise[1]
    want = (-(z**2) / 2).mean((1, 2, 3)) - torch.log(s_fix[0]) - 0.5 * math.log(2 * math.pi)
    assert torch.allclose(d1["logp"][0], want, atol=2e-5)


def test_edm_schedule_matches_reference():
    g = golden("edm_in64_small_T4_B2.npz")
    sched = samplers.edm_schedule(4)
    assert np.array_equal(sched["sigmas"].numpy(), g["sigmas"])
    assert np.array_equal(sched["sigma_up"].numpy(), g["sigma_up"])
    assert np.array_equal(sched["sigma_down"].numpy(), g["sigma_down"])
    assert np.array_equal(sched["log_betas_init"].numpy(), g["log_betas"])
    from diffusion_by_maxentirl_b200.schedule import EdmSchedule

    s = EdmSchedule(4)
    assert np.array_equal(s.sigmas.numpy(), g["sigmas"])
    assert np.array_equal(s.sigma_up.numpy(), g["sigma_up"]) and np.array_equal(s.sigma_down.numpy(), g["sigma_down"])
    # SURVEY App. B constants (T=10, rho=7) and the stochastic_last / rho=4 LSUN variant
    s10 = samplers.edm_schedule(10)
    assert np.allclose(s10["sigmas"][:4].numpy(), [80.0, 42.4152, 21.1087, 9.72320], rtol=1e-5)
    assert float(s10["sigma_up"][-1]) == 0.0 and float(s10["sigma_down"][-1]) == 0.0
    s4 = samplers.edm_schedule(4, rho=4.0, stochastic_last=True)
    assert np.allclose(s4["sigmas"].numpy(), [80.0, 27.7847, 6.57141, 0.674606, 0.002], rtol=1e-5)


EDM_SMALL = dict(image_size=32, model_channels=64, channel_mult=(1, 2, 3, 4), num_res_blocks=1, attention_ds=(2, 4, 8),
                 num_head_channels=64, use_scale_shift_norm=True)


def edm_small_weights(fp16_torso):
    g = golden("edm_in64_small_T4_B2.npz")
    keys = [str(k) for k in g["state_dict_keys"]]
    dtypes = [str(d) for d in g["state_dict_dtypes"]]
    shapes = edm_small_shapes()
    sd = synth.synth_state_dict({k: shapes[k] for k in keys if k != "log_betas"})
    if fp16_torso:
        for k, dt in zip(keys, dtypes):
            if k != "log_betas" and dt == "torch.float16":
                sd[k] = sd[k].half()
    sd["log_betas"] = torch.from_numpy(g["log_betas"])
    return sd, g


def edm_small_shapes():
    """state_dict shapes of the reduced-width ADM U-Net, derived from the oracle's own layout walk."""
    mc, ted = 64, 256
    shapes = {"time_embed.0.weight": (ted, mc), "time_embed.0.bias": (ted,), "time_embed.2.weight": (ted, ted),
              "time_embed.2.bias": (ted,), "label_emb.weight": (1000, ted)}
    inputs, outputs = nets.adm_layout(image_size=32, model_channels=mc, channel_mult=(1, 2, 3, 4), num_res_blocks=1,
                                      attention_ds=(2, 4, 8))

    def res(p, cin, cout):
        shapes.update({p + ".in_layers.0.weight": (cin,), p + ".in_layers.0.bias": (cin,),
                       p + ".in_layers.2.weight": (cout, cin, 3, 3), p + ".in_layers.2.bias": (cout,),
                       p + ".emb_layers.1.weight": (2 * cout, ted), p + ".emb_layers.1.bias": (2 * cout,),
                       p + ".out_layers.0.weight": (cout,), p + ".out_layers.0.bias": (cout,),
                       p + ".out_layers.3.weight": (cout, cout, 3, 3), p + ".out_layers.3.bias": (cout,)})
        if cin != cout:
            shapes.update({p + ".skip_connection.weight": (cout, cin, 1, 1), p + ".skip_connection.bias": (cout,)})

    def attn(p, c):
        shapes.update({p + ".norm.weight": (c,), p + ".norm.bias": (c,), p + ".qkv.weight": (3 * c, c, 1),
                       p + ".qkv.bias": (3 * c,), p + ".proj_out.weight": (c, c, 1), p + ".proj_out.bias": (c,)})

    def walk(prefix, layers):
        for j, (kind, cin, cout) in enumerate(layers):
            p = f"{prefix}.{j}"
            if kind == "conv":
                shapes.update({p + ".weight": (cout, 3, 3, 3), p + ".bias": (cout,)})
            elif kind == "attn":
                attn(p, cout)
            else:
                res(p, cin, cout)

    for i, layers in enumerate(inputs):
        walk(f"input_blocks.{i}", layers)
    walk("middle_block", [("res", 256, 256), ("attn", 256, 256), ("res", 256, 256)])
    for i, layers in enumerate(outputs):
        walk(f"output_blocks.{i}", layers)
    shapes.update({"out.0.weight": (64,), "out.0.bias": (64,), "out.2.weight": (3, 64, 3, 3), "out.2.bias": (3,)})
    return shapes


def test_edm_state_dict_layout_matches_reference():
    g = golden("edm_in64_small_T4_B2.npz")
    keys = [str(k) for k in g["state_dict_keys"]]
    shapes = edm_small_shapes()
    assert set(keys) - {"log_betas"} == set(shapes)
    # convert_to_fp16 (cm/unet.py:745-751): only convs inside input/middle/output blocks are fp16
    for k, dt in zip(keys, g["state_dict_dtypes"]):
        torso_conv = k.split(".")[0] in ("input_blocks", "middle_block", "output_blocks") and \
            (len(shapes.get(k, ())) >= 3 or (k.endswith(".bias") and len(shapes.get(k[:-4] + "weight", ())) >= 3))
        assert (str(dt) == "torch.float16") == torso_conv, (k, dt)


def test_oracle_edm_rollout_reproduces_reference():
    """fp16-torso oracle == the reference's only supported mode; fp32 oracle == stored fp32 states."""
    sd16, g = edm_small_weights(fp16_torso=True)
    # qkv / proj_out are Conv1d in the reference ([3C, C, 1]); the functional oracle applies them as 1x1 conv2d
    sd16 = {k: (v[..., None] if v.dim() == 3 else v) for k, v in sd16.items()}
    sched = samplers.edm_schedule(4)
    noise = synth.synth_noise(4, 2, (3, 32, 32))
    noise[0] = noise[0] * 80.0
    y = torch.from_numpy(g["y"])
    assert torch.equal(y, synth.synth_labels(2))
    with torch.no_grad():
        d = samplers.edm_rollout(lambda x, t, yy: nets.adm_unet_forward(sd16, x, t, yy, fp16_torso=True, **EDM_SMALL),
                                 sched, sd16["log_betas"], noise, y)
    for i in range(5):
        assert rel_l2(d["l_sample"][i], torch.from_numpy(g["l_sample_ref_fp16"][i])) < 1e-5
    assert d["sigma"][0].shape == (2,)  # [B], not [B,1,1,1] (openai_diffusion.py:96)
    sd32, _ = edm_small_weights(fp16_torso=False)
    sd32 = {k: (v[..., None] if v.dim() == 3 else v) for k, v in sd32.items()}
    with torch.no_grad():
        d32 = samplers.edm_rollout(lambda x, t, yy: nets.adm_unet_forward(sd32, x, t, yy, fp16_torso=False, **EDM_SMALL),
                                   sched, sd32["log_betas"], noise, y)
    for i in range(5):
        assert rel_l2(d32["l_sample"][i], torch.from_numpy(g["l_sample_fp32"][i])) < 1e-5
    assert rel_l2(d32["F"][0], torch.from_numpy(g["F_first_fp32"])) < 1e-5


def test_edm_noise_sigma_modes():
    sched = samplers.edm_schedule(10)
    lb = sched["log_betas_init"] + 0.1
    assert float(samplers.edm_noise_sigma(sched, lb, 9, 10, "fix_last")) == 0.0  # deterministic last step
    assert torch.allclose(samplers.edm_noise_sigma(sched, lb, 3, 10, "fix_last"), torch.exp(lb[3]))
    assert torch.equal(samplers.edm_noise_sigma(sched, lb, 7, 10, "fix_last3"), sched["sigma_up"][7])
    assert torch.equal(samplers.edm_noise_sigma(sched, lb, 2, 10, False), sched["sigma_up"][2])


def test_training_mode_oracle_reproduces_reference_gradients():
    """Row a9: tests/golden/ddpm_train_B2.npz holds eps and gradient slices of the *reference* unet_small.Model in train() mode
    (dropout 0.3 replaced by host-supplied masks, oracle/gen_golden.py::gen_ddpm_train, where the full comparison was bit-exact).
    The oracle with the same masks + torch autograd must reproduce them - this is the yardstick of the CUDA backward tests."""
    import json
    import os

    import numpy as np

    from oracle import nets, synth
    from oracle.gen_golden import ddpm_resblock_order, train_dropout_masks

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddpm_train_B2.npz"))
    shapes = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddpm_shapes.json")))["net"]
    sd = synth.synth_state_dict({k: tuple(v) for k, v in shapes.items()})
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    B, p_drop = 2, float(gold["p_drop"])
    cm, res = {"0": 128, "1": 256, "2": 256, "3": 256}, {"0": 32, "1": 16, "2": 8, "3": 4}
    mshapes = {}
    for k in ddpm_resblock_order():
        lvl = k.split(".")[1] if not k.startswith("mid") else "3"
        mshapes[k] = (B, cm[lvl], res[lvl], res[lvl])
    masks = train_dropout_masks(mshapes, p_drop, seed=77)
    g = torch.Generator().manual_seed(78)
    x = torch.randn(B, 3, 32, 32, generator=g)
    t = torch.tensor([394.07648, 66.86534])
    coef = torch.randn(B, 3, 32, 32, generator=g)
    out = nets.ddpm_unet_forward(sd, x, t, dropout_masks=masks)
    (out * coef).sum().backward()
    assert torch.allclose(out.detach(), torch.from_numpy(gold["eps"]), rtol=1e-5, atol=1e-5)
    n = 0
    for key in gold.files:
        if key.startswith("grad:"):
            ref = torch.from_numpy(gold[key])
            got = sd[key[5:]].grad[:8]
            err = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
            assert err < 1e-4, (key, err)
            n += 1
    assert n >= 8


def test_value_net_oracle_reproduces_reference_gradients():
    """tests/golden/value_train_B4.npz: parameter-gradient slices and the input gradient of the *reference* TimeIndependentValue under
    autograd (oracle/gen_golden.py::gen_ddpm_train, bit-exact there); the oracle must reproduce them without /root/reference."""
    import json
    import os

    import numpy as np

    from oracle import nets, synth

    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "golden", "value_train_B4.npz"))
    shapes = json.load(open(os.path.join(here, "golden", "ddpm_shapes.json")))["value"]
    sd = {k: v.clone().requires_grad_(True) for k, v in synth.synth_state_dict({k: tuple(v) for k, v in shapes.items()}, seed=1).items()}
    x = torch.from_numpy(gold["x"]).requires_grad_(True)
    (nets.value_forward(sd, x).flatten() * torch.from_numpy(gold["coef"])).sum().backward()
    assert torch.allclose(x.grad, torch.from_numpy(gold["dx"]), rtol=1e-4, atol=1e-7)
    for key in gold.files:
        if key.startswith("grad:"):
            ref = torch.from_numpy(gold[key])
            err = float((sd[key[5:]].grad[:8] - ref).norm() / ref.norm().clamp_min(1e-30))
            assert err < 1e-4, (key, err)


def test_adm_oracle_reproduces_reference_gradients():
    """Row f4: tests/golden/adm_train_B2.npz holds F and gradient slices of the *reference* UNetModel (reduced width, fp16 torso -
    the only mode it supports) in train() mode under autograd (oracle/gen_golden.py::gen_adm_train, where the full comparison
    was bit-exact).  The fp16-torso oracle + torch autograd must reproduce them without /root/reference; the fp32 oracle - the
    yardstick of tests/test_adm_train_gpu.py - must agree within the fp16 noise."""
    import numpy as np

    gold = golden("adm_train_B2.npz")
    shapes = edm_small_shapes()
    sd32 = synth.synth_state_dict(shapes)
    x, coef = torch.from_numpy(gold["x"]), torch.from_numpy(gold["coef"])
    t, y = torch.from_numpy(gold["t"]), torch.from_numpy(gold["y"])
    torso = ("input_blocks", "middle_block", "output_blocks")
    for fp16 in (True, False):
        sd = {}
        for k, v in sd32.items():
            v = v[..., None] if v.dim() == 3 else v
            # convert_to_fp16 (models/cm/unet.py:745-751): Conv weights / biases of the torso only
            is_conv = v.dim() == 4 or (k.endswith(".bias") and sd32[k[:-5] + ".weight"].dim() >= 3)
            if fp16 and k.split(".")[0] in torso and is_conv:
                v = v.half()
            sd[k] = v.clone().requires_grad_(True)
        out = nets.adm_unet_forward(sd, x, t, y, fp16_torso=fp16, **EDM_SMALL)
        (out * coef).sum().backward()
        ref = torch.from_numpy(gold["F"])
        e = float((out.detach().float() - ref).norm() / ref.norm())
        assert e < (1e-5 if fp16 else 1e-2), e
        n = 0
        for key in gold.files:
            if not key.startswith("grad:"):
                continue
            r = torch.from_numpy(gold[key])
            got = sd[key[5:]].grad.float().reshape(-1, *r.shape[1:])[:8].reshape(r.shape)
            err = float((got - r).norm() / r.norm().clamp_min(1e-30))
            assert err < (1e-4 if fp16 else 2e-2), (key, fp16, err)
            n += 1
        assert n >= 10
        rows = torch.from_numpy(gold["label_rows"])
        got = sd["label_emb.weight"].grad[y]
        assert float((got - rows).norm() / rows.norm()) < (1e-4 if fp16 else 2e-2)
