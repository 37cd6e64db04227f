"""CPU oracle for the IR-SDE / Denoising-SDE sampler hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch-CPU fp32 restatement of the
reference algorithm (Algolzw/image-restoration-sde @ 2598d73) for the one path this
repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package never does (it fails loudly when the CUDA library is missing instead).

Parity pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the *imported reference itself* (``tests/golden/make_golden.py`` ran
the reference in the build container and committed ``tests/golden/*.pt``);
``tests/test_oracle_golden.py`` replays them.

Every function cites the reference file:line it restates (paths relative to the
reference checkout, ``codes/...``).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# schedules  (codes/utils/sde_utils.py:84-152 IRSDE, :377-429 DenoisingSDE)
# --------------------------------------------------------------------------------------


def theta_schedule(T, schedule="cosine"):
    """theta[0..T] (length T+1).  sde_utils.py:91-121 (constant / linear / cosine)."""
    if schedule == "cosine":
        timesteps = T + 2
        steps = timesteps + 1
        x = torch.linspace(0, timesteps, steps, dtype=torch.float32)
        ac = torch.cos(((x / timesteps) + 0.008) / (1 + 0.008) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return 1 - ac[1:-1]
    if schedule == "linear":
        timesteps = T + 1
        scale = 1000 / timesteps
        return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float32)
    if schedule == "constant":
        return torch.ones(T + 1, dtype=torch.float32)
    raise ValueError("Not implemented such schedule yet!!!")


class Schedule:
    """thetas, sigmas, thetas_cumsum, sigma_bars, dt, max_sigma.

    sde_utils.py:86 (max_sigma/255 if >=1 for IRSDE), :379 (if >1 for DenoisingSDE),
    :141-144 (sigmas, cumsum - theta0, dt = -log(eps)/Theta_T, sigma_bars).
    """

    def __init__(self, max_sigma, T, schedule="cosine", eps=0.01, kind="irsde"):
        if kind == "irsde":
            self.max_sigma = max_sigma / 255 if max_sigma >= 1 else max_sigma
        else:  # DenoisingSDE: strict >, eps fixed to 0.04 (sde_utils.py:379,382)
            self.max_sigma = max_sigma / 255 if max_sigma > 1 else max_sigma
            eps = 0.04
            if schedule not in ("cosine",):
                schedule = "linear"
        self.T = T
        self.kind = kind
        thetas = theta_schedule(T, schedule)
        self.thetas = thetas
        self.sigmas = torch.sqrt(self.max_sigma ** 2 * 2 * thetas)
        self.thetas_cumsum = torch.cumsum(thetas, dim=0) - thetas[0]
        self.dt = -1 / self.thetas_cumsum[-1] * math.log(eps)
        self.sigma_bars = torch.sqrt(self.max_sigma ** 2 * (1 - torch.exp(-2 * self.thetas_cumsum * self.dt)))


# --------------------------------------------------------------------------------------
# sampler steps (same op order as the reference so fp32 rounding matches)
# --------------------------------------------------------------------------------------


def irsde_sde_step(s, x, mu, noise, z, t):
    """sde_utils.py:44-45,175-176,181-185: x - drift - dispersion, score = -noise/sigma_bar."""
    score = -noise / s.sigma_bars[t]
    drift = (s.thetas[t] * (mu - x) - s.sigmas[t] ** 2 * score) * s.dt
    disp = s.sigmas[t] * (z * math.sqrt(s.dt))
    return x - drift - disp


def irsde_ode_step(s, x, mu, noise, t):
    """sde_utils.py:47-48,178-179."""
    score = -noise / s.sigma_bars[t]
    drift = (s.thetas[t] * (mu - x) - 0.5 * s.sigmas[t] ** 2 * score) * s.dt
    return x - drift


def irsde_posterior_step(s, x, mu, noise, z, t):
    """sde_utils.py:197-223,237-239."""
    A0 = torch.exp(s.thetas_cumsum[t] * s.dt)
    x0 = (x - mu - s.sigma_bars[t] * noise) * A0 + mu
    A = torch.exp(-s.thetas[t] * s.dt)
    B = torch.exp(-s.thetas_cumsum[t] * s.dt)
    C = torch.exp(-s.thetas_cumsum[t - 1] * s.dt)
    term1 = A * (1 - C ** 2) / (1 - B ** 2)
    term2 = C * (1 - A ** 2) / (1 - B ** 2)
    mean = term1 * (x - mu) + term2 * (x0 - mu) + mu
    A2 = torch.exp(-2 * s.thetas[t] * s.dt)
    B2 = torch.exp(-2 * s.thetas_cumsum[t] * s.dt)
    C2 = torch.exp(-2 * s.thetas_cumsum[t - 1] * s.dt)
    var = (1 - A2) * (1 - C2) / (1 - B2)
    logvar = torch.log(torch.clamp(var, min=1e-20 * s.dt))
    std = (0.5 * logvar).exp() * s.max_sigma
    return mean + std * z


def dsde_sde_step(s, x, noise, z, t):
    """DenoisingSDE: sde_utils.py:450-452,458-462 (no mean-reversion drift)."""
    score = -noise / s.sigma_bars[t]
    A = torch.exp(-2 * s.thetas_cumsum[t] * s.dt)
    drift = -0.5 * s.sigmas[t] ** 2 * (1 + A) * score * s.dt
    disp = s.sigmas[t] * (z * math.sqrt(s.dt))
    return x - drift - disp


def dsde_ode_step(s, x, noise, t):
    """sde_utils.py:454-456."""
    score = -noise / s.sigma_bars[t]
    A = torch.exp(-2 * s.thetas_cumsum[t] * s.dt)
    drift = -0.5 * s.sigmas[t] ** 2 * A * score * s.dt
    return x - drift


def get_optimal_timestep(s, sigma, eps=1e-6):
    """sde_utils.py:550-554 (integer result, bit-exact)."""
    sigma = sigma / 255 if sigma > 1 else sigma
    hat = -1 / (2 * s.dt) * math.log(1 - sigma ** 2 / s.max_sigma ** 2 + eps)
    return int(torch.argmin((s.thetas_cumsum - hat).abs()))


def reverse_chain(s, net_fn, xT, mu, zs, mode="sde", T=-1):
    """Full chain, loop order reversed(range(1, T+1)) (sde_utils.py:252-299, :483-522).

    ``zs[i]`` is the pre-drawn N(0,1) tensor for the i-th executed step (t = T - i);
    the reference draws ``torch.randn_like`` in exactly this order.
    ``net_fn(x, t)`` returns the network's noise prediction.
    """
    T = s.T if T < 0 else T
    x = xT.clone()
    for i, t in enumerate(reversed(range(1, T + 1))):
        noise = net_fn(x, t)
        if mode == "sde":
            x = irsde_sde_step(s, x, mu, noise, zs[i], t)
        elif mode == "ode":
            x = irsde_ode_step(s, x, mu, noise, t)
        elif mode == "posterior":
            x = irsde_posterior_step(s, x, mu, noise, zs[i], t)
        elif mode == "dsde_sde":
            x = dsde_sde_step(s, x, noise, zs[i], t)
        elif mode == "dsde_ode":
            x = dsde_ode_step(s, x, noise, t)
        else:
            raise ValueError(mode)
    return x


# --------------------------------------------------------------------------------------
# score network: ConditionalUNet (functional; weights = reference state_dict)
# --------------------------------------------------------------------------------------


def _silu(x):
    return x * torch.sigmoid(x)


def time_embedding(P, time, nf):
    """SinusoidalPosEmb + time_mlp.  module_util.py:29-41, DenoisingUNet_arch.py:42-47."""
    half = nf // 2
    e = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half) * -e)
    emb = time[:, None] * freqs[None, :]
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    h = F.linear(emb, P["time_mlp.1.weight"], P["time_mlp.1.bias"])
    h = F.gelu(h)
    return F.linear(h, P["time_mlp.3.weight"], P["time_mlp.3.bias"])


def layer_norm_c(x, g):
    """Channel LayerNorm, eps=1e-5 for fp32.  module_util.py:70-79."""
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def _tr(trace, key, val):
    """Per-layer checkpoints for the GPU parity report (tests/): keyed by the state-dict name of the weight whose
    (fused) output this is, the same key the CUDA launch plan labels its ops with."""
    if trace is not None:
        trace[key] = val
    return val


def res_block(P, pre, x, temb, trace=None):
    """ResBlock / Block.  module_util.py:108-146 (convs bias-free, res_conv 1x1 or identity)."""
    ss = F.linear(_silu(temb), P[pre + "mlp.1.weight"], P[pre + "mlp.1.bias"])
    scale, shift = ss[:, :, None, None].chunk(2, dim=1)
    h = F.conv2d(x, P[pre + "block1.proj.weight"], padding=1)
    h = _tr(trace, pre + "block1.proj.weight", _silu(h * (scale + 1) + shift))
    h = _silu(F.conv2d(h, P[pre + "block2.proj.weight"], padding=1))
    if (pre + "res_conv.weight") in P:
        x = _tr(trace, pre + "res_conv.weight", F.conv2d(x, P[pre + "res_conv.weight"]))
    return _tr(trace, pre + "block2.proj.weight", h + x)


def linear_attention(P, pre, x, heads=4, dim_head=32, trace=None):
    """Residual(PreNorm(LinearAttention)).  module_util.py:20-26,82-90,150-178."""
    b, c, hh, ww = x.shape
    base = pre[:-3]  # "downs.0.2.fn." -> "downs.0.2."
    xn = _tr(trace, base + "norm", layer_norm_c(x, P[pre + "norm.g"]))
    qkv = F.conv2d(xn, P[pre + "fn.to_qkv.weight"])
    q, k, v = [t.reshape(b, heads, dim_head, hh * ww) for t in qkv.chunk(3, dim=1)]
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    v = v / (hh * ww)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, heads * dim_head, hh, ww)
    out = _tr(trace, pre + "fn.to_out.0.weight", F.conv2d(out, P[pre + "fn.to_out.0.weight"], P[pre + "fn.to_out.0.bias"]))
    out = layer_norm_c(out, P[pre + "fn.to_out.1.g"])
    return _tr(trace, base + "to_out.norm+res", out + x)


def full_attention(P, pre, x, heads=4, dim_head=32, trace=None):
    """Residual(PreNorm(Attention)) - denoising-sde mid_attn only.  module_util.py:182-204."""
    b, c, hh, ww = x.shape
    base = pre[:-3]
    xn = _tr(trace, base + "norm", layer_norm_c(x, P[pre + "norm.g"]))
    qkv = F.conv2d(xn, P[pre + "fn.to_qkv.weight"])
    q, k, v = [t.reshape(b, heads, dim_head, hh * ww) for t in qkv.chunk(3, dim=1)]
    q = q * dim_head ** -0.5
    sim = torch.einsum("bhdi,bhdj->bhij", q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)
    out = _tr(trace, base + "full attention", out.permute(0, 1, 3, 2).reshape(b, heads * dim_head, hh, ww))
    out = F.conv2d(out, P[pre + "fn.to_out.weight"], P[pre + "fn.to_out.bias"])
    return _tr(trace, pre + "fn.to_out.weight", out + x)


def unet_forward(P, xt, cond, time, nf, depth, variant="conditional", trace=None):
    # `depth` may be the latent-task variant's ch_mult list (latent-dehazing/.../DenoisingUNet_arch.py:20): the forward
    # is the same, only the number of levels (= len(ch_mult)) matters here - channel widths come from the weights
    if isinstance(depth, (list, tuple)):
        depth = len(depth)
    """ConditionalUNet.forward.  codes/config/deraining/models/modules/DenoisingUNet_arch.py:85-134
    (variant="conditional": input cat(xt-cond, cond)); the denoising-sde variant
    (codes/config/denoising-sde/models/modules/DenoisingUNet_arch.py:84-132) takes x only and
    uses full Attention at mid_attn.
    """
    if isinstance(time, (int, float)):
        time = torch.tensor([time])
    time = time.reshape(-1)
    if variant == "conditional":
        x = torch.cat([xt - cond, cond], dim=1)
    else:
        x = xt
    H, W = x.shape[2:]
    s = 2 ** depth
    x = F.pad(x, (0, (s - W % s) % s, 0, (s - H % s) % s), "reflect")
    x = _tr(trace, "init_conv.weight", F.conv2d(x, P["init_conv.weight"], padding=3))
    x_ = x
    temb = time_embedding(P, time.to(torch.float32) if time.dtype.is_floating_point else time, nf)
    h = []
    for i in range(depth):
        pre = "downs.%d." % i
        x = res_block(P, pre + "0.", x, temb, trace)
        h.append(x)
        x = res_block(P, pre + "1.", x, temb, trace)
        x = linear_attention(P, pre + "2.fn.", x, trace=trace)
        h.append(x)
        if i != depth - 1:
            x = _tr(trace, pre + "3.weight", F.conv2d(x, P[pre + "3.weight"], P[pre + "3.bias"], stride=2, padding=1))
        else:
            x = _tr(trace, pre + "3.weight", F.conv2d(x, P[pre + "3.weight"], padding=1))
    x = res_block(P, "mid_block1.", x, temb, trace)
    if variant == "conditional":
        x = linear_attention(P, "mid_attn.fn.", x, trace=trace)
    else:
        x = full_attention(P, "mid_attn.fn.", x, trace=trace)
    x = res_block(P, "mid_block2.", x, temb, trace)
    for j in range(depth):
        i = depth - 1 - j
        pre = "ups.%d." % j
        x = torch.cat([x, h.pop()], dim=1)
        x = res_block(P, pre + "0.", x, temb, trace)
        x = torch.cat([x, h.pop()], dim=1)
        x = res_block(P, pre + "1.", x, temb, trace)
        x = linear_attention(P, pre + "2.fn.", x, trace=trace)
        if i != 0:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = _tr(trace, pre + "3.1.weight", F.conv2d(x, P[pre + "3.1.weight"], P[pre + "3.1.bias"], padding=1))
        else:
            x = _tr(trace, pre + "3.weight", F.conv2d(x, P[pre + "3.weight"], padding=1))
    x = torch.cat([x, x_], dim=1)
    x = res_block(P, "final_res_block.", x, temb, trace)
    x = F.conv2d(x, P["final_conv.weight"], P["final_conv.bias"], padding=1)
    return x[..., :H, :W]


# --------------------------------------------------------------------------------------
# deterministic synthetic weights with the reference's state-dict names/shapes
# --------------------------------------------------------------------------------------


def unet_param_shapes(in_nc, out_nc, nf, depth, variant="conditional"):
    """State-dict names -> shapes, in the reference's registration order
    (DenoisingUNet_arch.py:19-76, module_util.py:125-161,185-190).  `depth` may be a ch_mult list: the latent-task
    variant (latent-dehazing/.../DenoisingUNet_arch.py:20,51-56,70) whose level i has nf*[1,ch_mult...][i] channels."""
    mult = [1] + list(depth) if isinstance(depth, (list, tuple)) else [2 ** i for i in range(depth + 1)]
    depth = len(mult) - 1
    S = {}
    td = nf * 4
    cin0 = in_nc * 2 if variant == "conditional" else in_nc
    S["init_conv.weight"] = (nf, cin0, 7, 7)
    S["time_mlp.1.weight"] = (td, nf)
    S["time_mlp.1.bias"] = (td,)
    S["time_mlp.3.weight"] = (td, td)
    S["time_mlp.3.bias"] = (td,)

    def rb(pre, ci, co):
        S[pre + "mlp.1.weight"] = (2 * co, td)
        S[pre + "mlp.1.bias"] = (2 * co,)
        S[pre + "block1.proj.weight"] = (co, ci, 3, 3)
        S[pre + "block2.proj.weight"] = (co, co, 3, 3)
        if ci != co:
            S[pre + "res_conv.weight"] = (co, ci, 1, 1)

    def la(pre, c, full=False):
        S[pre + "fn.fn.to_qkv.weight"] = (384, c, 1, 1)
        if full:
            S[pre + "fn.fn.to_out.weight"] = (c, 128, 1, 1)
            S[pre + "fn.fn.to_out.bias"] = (c,)
        else:
            S[pre + "fn.fn.to_out.0.weight"] = (c, 128, 1, 1)
            S[pre + "fn.fn.to_out.0.bias"] = (c,)
            S[pre + "fn.fn.to_out.1.g"] = (1, c, 1, 1)
        S[pre + "fn.norm.g"] = (1, c, 1, 1)

    ups = []
    for i in range(depth):
        di, do = nf * mult[i], nf * mult[i + 1]
        pre = "downs.%d." % i
        rb(pre + "0.", di, di)
        rb(pre + "1.", di, di)
        la(pre + "2.", di)
        if i != depth - 1:
            S[pre + "3.weight"] = (do, di, 4, 4)
            S[pre + "3.bias"] = (do,)
        else:
            S[pre + "3.weight"] = (do, di, 3, 3)
        ups.insert(0, (i, di, do))
    for j, (i, di, do) in enumerate(ups):
        pre = "ups.%d." % j
        rb(pre + "0.", do + di, do)
        rb(pre + "1.", do + di, do)
        la(pre + "2.", do)
        if i != 0:
            S[pre + "3.1.weight"] = (di, do, 3, 3)
            S[pre + "3.1.bias"] = (di,)
        else:
            S[pre + "3.weight"] = (di, do, 3, 3)
    mid = nf * mult[depth]
    rb("mid_block1.", mid, mid)
    la("mid_attn.", mid, full=(variant != "conditional"))
    rb("mid_block2.", mid, mid)
    rb("final_res_block.", 2 * nf, nf)
    S["final_conv.weight"] = (out_nc, nf, 3, 3)
    S["final_conv.bias"] = (out_nc,)
    return S


def make_weights(in_nc, out_nc, nf, depth, variant="conditional", seed=0, out_gain=1.0):
    """Synthetic weights (kaiming-uniform-like fan-in scaling, non-trivial biases / gains).
    Not the reference initialiser: parity tests feed the SAME dict to oracle and CUDA."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in unet_param_shapes(in_nc, out_nc, nf, depth, variant).items():
        if name.endswith(".g"):
            P[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("bias"):
            P[name] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            bound = math.sqrt(3.0 / fan_in)
            P[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    P["final_conv.weight"] *= out_gain
    return P


# --------------------------------------------------------------------------------------
# score network #2: ConditionalNAFNet (Refusion)  codes/config/deraining/models/modules/DenoisingNAFNet_arch.py
# --------------------------------------------------------------------------------------


def _gate(x, dim=1):
    """SimpleGate: DenoisingNAFNet_arch.py:9-12."""
    a, b = x.chunk(2, dim=dim)
    return a * b


def naf_block(P, pre, x, t):
    """NAFBlock.forward: DenoisingNAFNet_arch.py:51-84 (time chunk order: shift_att, scale_att, shift_ffn, scale_ffn)."""
    te = F.linear(_gate(t, dim=-1), P[pre + "mlp.1.weight"], P[pre + "mlp.1.bias"])[:, :, None, None]
    shift_att, scale_att, shift_ffn, scale_ffn = te.chunk(4, dim=1)
    inp = x
    x = layer_norm_c(x, P[pre + "norm1.g"])
    x = x * (scale_att + 1) + shift_att
    x = F.conv2d(x, P[pre + "conv1.weight"], P[pre + "conv1.bias"])
    x = F.conv2d(x, P[pre + "conv2.weight"], P[pre + "conv2.bias"], padding=1, groups=x.shape[1])
    x = _gate(x)
    pooled = x.mean(dim=(2, 3), keepdim=True)
    x = x * F.conv2d(pooled, P[pre + "sca.1.weight"], P[pre + "sca.1.bias"])
    x = F.conv2d(x, P[pre + "conv3.weight"], P[pre + "conv3.bias"])
    y = inp + x * P[pre + "beta"]
    x = layer_norm_c(y, P[pre + "norm2.g"])
    x = x * (scale_ffn + 1) + shift_ffn
    x = F.conv2d(x, P[pre + "conv4.weight"], P[pre + "conv4.bias"])
    x = _gate(x)
    x = F.conv2d(x, P[pre + "conv5.weight"], P[pre + "conv5.bias"])
    return y + x * P[pre + "gamma"]


def nafnet_forward(P, inp, cond, time, width, enc_blk_nums, middle_blk_num, dec_blk_nums, latent=False):
    """ConditionalNAFNet.forward: DenoisingNAFNet_arch.py:147-188; latent variant
    (codes/config/latent-dehazing/models/modules/DenoisingNAFNet_arch.py:147-181): ending(x + encs[0])."""
    if isinstance(time, (int, float)):
        time = torch.tensor([time])
    time = time.reshape(-1)
    x = torch.cat([inp - cond, cond], dim=1)
    half = width // 2
    freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    emb = time[:, None] * freqs[None, :]
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    t = F.linear(emb, P["time_mlp.1.weight"], P["time_mlp.1.bias"])
    t = F.linear(_gate(t, dim=-1), P["time_mlp.3.weight"], P["time_mlp.3.bias"])
    H, W = x.shape[2:]
    s = 2 ** len(enc_blk_nums)
    x = F.pad(x, (0, (s - W % s) % s, 0, (s - H % s) % s))  # zero pad (not reflect)
    x = F.conv2d(x, P["intro.weight"], P["intro.bias"], padding=1)
    x_intro = x
    encs = []
    for i, num in enumerate(enc_blk_nums):
        for j in range(num):
            x = naf_block(P, "encoders.%d.%d." % (i, j), x, t)
        encs.append(x)
        x = F.conv2d(x, P["downs.%d.weight" % i], P["downs.%d.bias" % i], stride=2)
    for j in range(middle_blk_num):
        x = naf_block(P, "middle_blks.%d." % j, x, t)
    for i, num in enumerate(dec_blk_nums):
        x = F.pixel_shuffle(F.conv2d(x, P["ups.%d.0.weight" % i]), 2)
        x = x + encs[len(encs) - 1 - i]
        for j in range(num):
            x = naf_block(P, "decoders.%d.%d." % (i, j), x, t)
    if latent:
        x = x + x_intro
    x = F.conv2d(x, P["ending.weight"], P["ending.bias"], padding=1)
    return x[..., :H, :W]


def nafnet_param_shapes(img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums):
    """State-dict names -> shapes in the reference's registration order (DenoisingNAFNet_arch.py:15-49,89-143)."""
    S = {}
    td = width * 4
    S["time_mlp.1.weight"] = (td * 2, width)
    S["time_mlp.1.bias"] = (td * 2,)
    S["time_mlp.3.weight"] = (td, td)
    S["time_mlp.3.bias"] = (td,)
    S["intro.weight"] = (width, img_channel * 2, 3, 3)
    S["intro.bias"] = (width,)
    S["ending.weight"] = (img_channel, width, 3, 3)
    S["ending.bias"] = (img_channel,)

    def blk(pre, c):
        S[pre + "beta"] = (1, c, 1, 1)
        S[pre + "gamma"] = (1, c, 1, 1)
        S[pre + "mlp.1.weight"] = (4 * c, td // 2)
        S[pre + "mlp.1.bias"] = (4 * c,)
        S[pre + "conv1.weight"] = (2 * c, c, 1, 1)
        S[pre + "conv1.bias"] = (2 * c,)
        S[pre + "conv2.weight"] = (2 * c, 1, 3, 3)
        S[pre + "conv2.bias"] = (2 * c,)
        S[pre + "conv3.weight"] = (c, c, 1, 1)
        S[pre + "conv3.bias"] = (c,)
        S[pre + "sca.1.weight"] = (c, c, 1, 1)
        S[pre + "sca.1.bias"] = (c,)
        S[pre + "conv4.weight"] = (2 * c, c, 1, 1)
        S[pre + "conv4.bias"] = (2 * c,)
        S[pre + "conv5.weight"] = (c, c, 1, 1)
        S[pre + "conv5.bias"] = (c,)
        S[pre + "norm1.g"] = (1, c, 1, 1)
        S[pre + "norm2.g"] = (1, c, 1, 1)

    chan = width
    enc_c = []
    for i, num in enumerate(enc_blk_nums):
        for j in range(num):
            blk("encoders.%d.%d." % (i, j), chan)
        enc_c.append(chan)
        chan *= 2
    mid_c = chan
    dec_c = []
    for i, num in enumerate(dec_blk_nums):
        chan //= 2
        for j in range(num):
            blk("decoders.%d.%d." % (i, j), chan)
        dec_c.append(chan)
    for j in range(middle_blk_num):
        blk("middle_blks.%d." % j, mid_c)
    chan = mid_c
    for i in range(len(dec_blk_nums)):
        S["ups.%d.0.weight" % i] = (chan * 2, chan, 1, 1)
        chan //= 2
    chan = width
    for i in range(len(enc_blk_nums)):
        S["downs.%d.weight" % i] = (chan * 2, chan, 2, 2)
        S["downs.%d.bias" % i] = (chan * 2,)
        chan *= 2
    return S


def make_nafnet_weights(img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums, seed=0):
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in nafnet_param_shapes(img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums).items():
        if name.endswith(".g"):
            P[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("beta") or name.endswith("gamma"):
            P[name] = 0.5 * torch.randn(shp, generator=g)  # reference inits these to zero (identity blocks)
        elif name.endswith("bias"):
            P[name] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            P[name] = (torch.rand(shp, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
    return P


# --------------------------------------------------------------------------------------
# Refusion latent autoencoder: UNet.encode / UNet.decode   codes/config/latent-dehazing/models/modules/UNet_arch.py:17-97
# --------------------------------------------------------------------------------------


def _res_block_notime(P, pre, x):
    """ResBlock with time_emb_dim=None (module_util.py:125-146: no scale/shift)."""
    h = _silu(F.conv2d(x, P[pre + "block1.proj.weight"], padding=1))
    h = _silu(F.conv2d(h, P[pre + "block2.proj.weight"], padding=1))
    if (pre + "res_conv.weight") in P:
        x = F.conv2d(x, P[pre + "res_conv.weight"])
    return h + x


def latent_unet_encode(P, x, ch_mult):
    """UNet.encode (UNet_arch.py:59-76): returns (z, h) with h the list of skip tensors."""
    depth = len(ch_mult)
    H, W = x.shape[2:]
    s = 2 ** depth
    x = F.pad(x, (0, (s - W % s) % s, 0, (s - H % s) % s), "reflect")
    x = F.conv2d(x, P["init_conv.weight"], padding=1)
    h = [x]
    for i in range(depth):
        pre = "encoder.%d." % i
        x = _res_block_notime(P, pre + "0.", x)
        h.append(x)
        x = _res_block_notime(P, pre + "1.", x)
        if i == depth - 1:
            x = linear_attention(P, pre + "2.fn.", x)
        h.append(x)
        if i != depth - 1:
            x = F.conv2d(x, P[pre + "3.weight"], P[pre + "3.bias"], stride=2, padding=1)
        else:
            x = F.conv2d(x, P[pre + "3.weight"], padding=1)
    return F.conv2d(x, P["latent_conv.weight"]), h


def latent_unet_decode(P, z, h, ch_mult, H, W):
    """UNet.decode (UNet_arch.py:78-91)."""
    depth = len(ch_mult)
    x = F.conv2d(z, P["post_latent_conv.weight"])
    for i in range(depth):
        pre = "decoder.%d." % i
        x = torch.cat([x, h[-(i * 2 + 1)]], dim=1)
        x = _res_block_notime(P, pre + "0.", x)
        x = torch.cat([x, h[-(i * 2 + 2)]], dim=1)
        x = _res_block_notime(P, pre + "1.", x)
        if i == 0:
            x = linear_attention(P, pre + "2.fn.", x)
        if i != depth - 1:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, P[pre + "3.1.weight"], P[pre + "3.1.bias"], padding=1)
        else:
            x = F.conv2d(x, P[pre + "3.weight"], padding=1)
    x = F.conv2d(x + h[0], P["final_conv.weight"], P["final_conv.bias"], padding=1)
    return x[..., :H, :W]


def latent_unet_param_shapes(in_ch, out_ch, ch, ch_mult, embed_dim):
    """State-dict names -> shapes in the reference's registration order (UNet_arch.py:18-51)."""
    S = {}
    depth = len(ch_mult)
    m = [1] + list(ch_mult)
    S["init_conv.weight"] = (ch, in_ch, 3, 3)

    def rb(pre, ci, co):
        S[pre + "block1.proj.weight"] = (co, ci, 3, 3)
        S[pre + "block2.proj.weight"] = (co, co, 3, 3)
        if ci != co:
            S[pre + "res_conv.weight"] = (co, ci, 1, 1)

    def la(pre, c):
        S[pre + "fn.fn.to_qkv.weight"] = (384, c, 1, 1)
        S[pre + "fn.fn.to_out.0.weight"] = (c, 128, 1, 1)
        S[pre + "fn.fn.to_out.0.bias"] = (c,)
        S[pre + "fn.fn.to_out.1.g"] = (1, c, 1, 1)
        S[pre + "fn.norm.g"] = (1, c, 1, 1)

    dec = []
    for i in range(depth):
        di, do = ch * m[i], ch * m[i + 1]
        pre = "encoder.%d." % i
        rb(pre + "0.", di, di)
        rb(pre + "1.", di, di)
        if i == depth - 1:
            la(pre + "2.", di)
        if i != depth - 1:
            S[pre + "3.weight"] = (do, di, 4, 4)
            S[pre + "3.bias"] = (do,)
        else:
            S[pre + "3.weight"] = (do, di, 3, 3)
        dec.insert(0, (i, di, do))
    for j, (i, di, do) in enumerate(dec):
        pre = "decoder.%d." % j
        rb(pre + "0.", do + di, do)
        rb(pre + "1.", do + di, do)
        if i == depth - 1:
            la(pre + "2.", do)
        if i != 0:
            S[pre + "3.1.weight"] = (di, do, 3, 3)
            S[pre + "3.1.bias"] = (di,)
        else:
            S[pre + "3.weight"] = (di, do, 3, 3)
    mid = ch * m[-1]
    S["latent_conv.weight"] = (embed_dim, mid, 1, 1)
    S["post_latent_conv.weight"] = (mid, embed_dim, 1, 1)
    S["final_conv.weight"] = (out_ch, ch, 3, 3)
    S["final_conv.bias"] = (out_ch,)
    return S


def make_latent_unet_weights(in_ch, out_ch, ch, ch_mult, embed_dim, seed=0):
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in latent_unet_param_shapes(in_ch, out_ch, ch, ch_mult, embed_dim).items():
        if name.endswith(".g"):
            P[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("bias"):
            P[name] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            P[name] = (torch.rand(shp, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
    return P
