"""Pin the CPU oracle against outputs of the imported reference (tests/golden/make_golden.py)."""
import torch

from oracle import irsde_oracle as O


def _close(a, b, tol):
    d = (a - b).abs().max().item()
    assert d <= tol, d


def test_irsde_schedules(golden):
    for g in golden["irsde_schedules"]:
        ms, T, s, eps = g["args"]
        sc = O.Schedule(ms, T, s, eps)
        for k in ("thetas", "sigmas", "thetas_cumsum", "sigma_bars"):
            assert torch.equal(getattr(sc, k), g[k]), k  # same torch ops => bit-exact
        assert torch.equal(torch.as_tensor(sc.dt), g["dt"])
        assert sc.max_sigma == g["max_sigma"]


def test_schedule_known_answers():
    # SURVEY.md 8(a-1) known answers measured on the reference
    sc = O.Schedule(10, 100, "cosine", 0.005)
    assert abs(float(sc.dt) - 0.10409380) < 1e-7
    assert abs(float(sc.thetas[1]) - 0.00169456) < 1e-7
    assert abs(float(sc.thetas_cumsum[100]) - 50.899448) < 1e-4
    assert abs(float(sc.sigma_bars[100]) - 0.03921520) < 1e-7
    assert abs(float(O.Schedule(10, 50, "cosine", 0.005).dt) - 0.20614207) < 1e-7


def test_dsde_schedules(golden):
    for g in golden["dsde_schedules"]:
        ms, T, s = g["args"]
        sc = O.Schedule(ms, T, s, kind="dsde")
        for k in ("thetas", "sigmas", "thetas_cumsum", "sigma_bars"):
            assert torch.equal(getattr(sc, k), g[k]), k
        for sg, t in g["optimal_t"].items():
            assert O.get_optimal_timestep(sc, sg) == t  # integer work: bit-exact


def test_irsde_steps(golden):
    g = golden["irsde_steps"]
    sc = O.Schedule(*g["args"])
    for st in g["steps"]:
        t, z = st["t"], st["z"]
        assert torch.equal(O.irsde_sde_step(sc, g["x"], g["mu"], g["noise"], z, t), st["sde"])
        assert torch.equal(O.irsde_ode_step(sc, g["x"], g["mu"], g["noise"], t), st["ode"])
        assert torch.equal(O.irsde_posterior_step(sc, g["x"], g["mu"], g["noise"], z, t), st["posterior"])


def test_dsde_steps(golden):
    g = golden["dsde_steps"]
    sc = O.Schedule(*g["args"], kind="dsde")
    for st in g["steps"]:
        assert torch.equal(O.dsde_sde_step(sc, g["x"], g["noise"], st["z"], st["t"]), st["sde"])
        assert torch.equal(O.dsde_ode_step(sc, g["x"], g["noise"], st["t"]), st["ode"])


def test_unet_conditional(golden):
    g = golden["unet_cond"]
    y = O.unet_forward(g["state"], g["xt"], g["cond"], g["t_int"], g["nf"], g["depth"])
    _close(y, g["y_int"], 1e-5)
    y = O.unet_forward(g["state"], g["xt"], g["cond"], g["t_vec"], g["nf"], g["depth"])
    _close(y, g["y_vec"], 1e-5)
    shapes = O.unet_param_shapes(3, 3, g["nf"], g["depth"])
    assert list(shapes) == list(g["state"])  # same names, same registration order
    assert all(tuple(g["state"][k].shape) == tuple(v) for k, v in shapes.items())


def test_unet_denoising_variant(golden):
    g = golden["unet_dsde"]
    y = O.unet_forward(g["state"], g["x"], None, g["t_int"], g["nf"], g["depth"], variant="denoising")
    _close(y, g["y"], 1e-5)
    shapes = O.unet_param_shapes(3, 3, g["nf"], g["depth"], variant="denoising")
    assert list(shapes) == list(g["state"])


def test_chains(golden):
    g = golden["irsde_chain"]
    u = golden["unet_cond"]
    sc = O.Schedule(*g["args"])
    net = lambda x, t: O.unet_forward(u["state"], x, g["lq"], t, u["nf"], u["depth"])
    for mode, c in g["chains"].items():
        x0 = O.reverse_chain(sc, net, g["xT"], g["lq"], c["zs"], mode)
        _close(x0, c["x0"], 2e-4)  # chain is numerically expansive; observed ~1e-6
        x5 = O.reverse_chain(sc, net, g["xT"], g["lq"], c["zs"], mode, T=5)
        _close(x5, c["x0_T5"], 2e-4)


def test_dsde_chain(golden):
    g = golden["unet_dsde"]
    sc = O.Schedule(*g["args"], kind="dsde")
    net = lambda x, t: O.unet_forward(g["state"], x, None, t, g["nf"], g["depth"], variant="denoising")
    assert O.get_optimal_timestep(sc, 25) == g["Tstar"]
    _close(O.reverse_chain(sc, net, g["x"], None, g["zs"], "dsde_sde", T=g["Tstar"]), g["x0_sde"], 2e-4)
    _close(O.reverse_chain(sc, net, g["x"], None, g["zs"], "dsde_ode", T=g["Tstar"]), g["x0_ode"], 2e-4)


def test_nafnet_oracle_vs_reference():
    import os
    g_all = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_nafnet.pt"), weights_only=False)
    for key, g in g_all.items():
        c = g["cfg"]
        args = (c["width"], c["enc_blk_nums"], c["middle_blk_num"], c["dec_blk_nums"])
        y = O.nafnet_forward(g["state"], g["x"], g["cond"], g["t_int"], *args, latent=g["latent"])
        _close(y, g["y"], 1e-5)
        yv = O.nafnet_forward(g["state"], g["x"], g["cond"], g["t_vec"], *args, latent=g["latent"])
        _close(yv, g["y_vec"], 1e-5)
        shapes = O.nafnet_param_shapes(c["img_channel"], c["width"], c["middle_blk_num"], c["enc_blk_nums"], c["dec_blk_nums"])
        assert list(shapes) == list(g["state"])
