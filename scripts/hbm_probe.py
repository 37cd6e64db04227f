"""Outputs of a fixed set of forwards under the IRSDE_HBM_NEW / IRSDE_LN_PP setting of the environment (round-2b HBM kernels).

  python scripts/hbm_probe.py run  OUT.pt     # one setting (the library reads the env once, at load)
  python scripts/hbm_probe.py cmp  REF.pt NEW.pt [...]   # NEW vs REF (REF = IRSDE_HBM_NEW=0, the round-2a kernels)

The probe only sorts out a kernel that is plainly wrong before GPU time is spent on the test suite (which compares with the
oracle): fp32 / fp32x3 outputs must agree to 1e-5 of the output range (LayerNorm / merge / fold keep their arithmetic there),
bf16 outputs to 3 % of the range (rsqrt / reciprocal forms and a different fp32 summation tree in the k/v pass move bf16
roundings)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def run(out):
    import irsde_b200
    dev = torch.device("cuda:0")
    res = {}
    g = torch.Generator().manual_seed(1234)
    def rnd(*s):
        return torch.rand(*s, generator=g).to(dev)
    cases = [("unet_bf16_256", "bf16", 64, (2, 3, 256, 256)),      # C = 64..1024: LPP 8/16/32, MAXV 1/2/4; 1 024 k/v chunks
             ("unet_bf16_ragged", "bf16", 64, (3, 3, 40, 40)),     # ragged pixel counts / ragged k/v chunks (1600, 400, 100, 25)
             ("unet_bf16_nf32", "bf16", 32, (1, 3, 72, 56)),
             ("unet_fp32_ragged", "fp32", 64, (1, 3, 40, 40)),     # fp32: LPP 16/32, MAXV 1/2/4/8
             ("unet_fp32x3", "fp32x3", 64, (1, 3, 64, 64))]
    attn_only = os.environ.get("HBM_PROBE_ATTN_ONLY") == "1"   # only the cases that run the LinearAttention kernels
    if attn_only:
        cases = [c for c in cases if c[0] in ("unet_bf16_256", "unet_bf16_ragged", "unet_fp32_ragged")]
    for name, prec, nf, shp in cases:
        torch.manual_seed(7)
        net = irsde_b200.ConditionalUNet(3, 3, nf, depth=4, precision=prec).to(dev)
        x, c = rnd(*shp), rnd(*shp)
        res[name] = net(x, c, 17).float().cpu()
        if name == "unet_bf16_ragged":   # the multi-GPU partition property: batch-sharded == unsharded, bit for bit
            rows = torch.cat([net(x[i:i + 1].contiguous(), c[i:i + 1].contiguous(), 17).float().cpu() for i in range(shp[0])])
            res["sharded_equal"] = torch.tensor(float(torch.equal(rows, res[name])))
        del net
    torch.manual_seed(8)
    net = irsde_b200.DenoisingUNet(3, 3, 64, depth=4, precision="bf16").to(dev)
    res["dunet_bf16"] = net(rnd(2, 3, 64, 64), 5).float().cpu()
    del net
    for prec in (() if attn_only else ("bf16", "fp32")):
        torch.manual_seed(9)
        net = irsde_b200.ConditionalNAFNet(img_channel=3, width=64, middle_blk_num=1, enc_blk_nums=[1, 1, 1], dec_blk_nums=[1, 1, 1],
                                           precision=prec).to(dev)   # LayerNorm with the time modulation rows
        with torch.no_grad():
            for n, p in net.named_parameters():
                if n.endswith("beta") or n.endswith("gamma"):   # zero-initialised in the reference: blocks would be identities
                    p.fill_(0.3)
        net.sync_weights(dev)
        x, c = rnd(2, 3, 40, 48), rnd(2, 3, 40, 48)
        res["naf_" + prec] = net(x, c, 11).float().cpu()
        del net
    torch.cuda.synchronize()
    torch.save(res, out)
    print("probe", os.environ.get("IRSDE_HBM_NEW"), os.environ.get("IRSDE_LN_PP"), {k: float(v.abs().max()) for k, v in res.items()})


def cmp(ref, new):
    a, b = torch.load(ref), torch.load(new)
    ok = True
    for k in a:
        if k == "sharded_equal":
            good = float(b[k]) == 1.0
            ok = ok and good
            print("  %-18s batch-sharded == unsharded bit for bit: %s" % (k, "ok" if good else "FAIL"))
            continue
        rng = float(a[k].abs().max()) + 1e-12
        d = float((a[k] - b[k]).abs().max()) / rng
        tol = float(os.environ.get("HBM_PROBE_TOL_BF16", "3e-2")) if "bf16" in k else 1e-5
        good = d <= tol and bool(torch.isfinite(b[k]).all())
        ok = ok and good
        print("  %-18s rel.max diff %.3e (tol %.0e) %s" % (k, d, tol, "ok" if good else "FAIL"))
    return ok


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        bad = [n for n in sys.argv[3:] if not cmp(sys.argv[2], n)]
        print("BAD", bad)
        sys.exit(1 if bad else 0)
