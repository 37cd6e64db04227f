"""2-GPU test of the C-ABI NCCL entry points (irsde_comm_init / irsde_broadcast_weights / irsde_gather): batch-sharded
reverse_sde over two processes == the single-GPU chain, bit for bit (SURVEY 8 b, e).  Skipped with fewer than 2 devices
(run with `gpurun --gpus 2`)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, tmpdir, B, precision):
    sys.path.insert(0, ROOT)
    import time
    import irsde_b200
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.manual_seed(100 + rank)           # DIFFERENT random weights per rank: only the broadcast can make them agree
    net = irsde_b200.ConditionalUNet(3, 3, 16, depth=2, precision=precision).to(dev)
    idf = os.path.join(tmpdir, "nccl_id")
    if rank == 0:
        uid = irsde_b200.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idf + ".tmp", idf)
    else:
        for _ in range(600):
            if os.path.exists(idf):
                break
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    comm = irsde_b200.NativeComm(net, uid, rank, world, dev)
    comm.broadcast_weights(src=0)
    sde = irsde_b200.IRSDE(10, 8, eps=0.005, device=dev)
    sde.set_model(net)
    g = torch.Generator().manual_seed(7)
    lq = torch.rand(B, 3, 24, 32, generator=g)
    xT = lq + torch.randn(B, 3, 24, 32, generator=g) * sde.max_sigma
    zs = torch.randn(8, B, 3, 24, 32, generator=g)
    lo, hi = irsde_b200.shard_range(B, rank, world)
    sde.set_mu(lq[lo:hi].to(dev))
    part = sde.reverse_sde(xT[lo:hi].to(dev), zs=zs[:, lo:hi].to(dev))
    full = comm.gather(part, B)
    # philox path: image_base = the rank's first global image
    sde.rng, sde.seed, sde.seed_auto_increment, sde.image_base = "philox", 3, False, lo
    part_p = sde.reverse_sde(xT[lo:hi].to(dev))
    full_p = comm.gather(part_p, B)
    torch.cuda.synchronize()
    torch.save({"full": full.cpu(), "full_p": full_p.cpu()}, os.path.join(tmpdir, "out%d.pt" % rank))
    if rank == 0:   # the single-GPU answer with rank 0's (= everyone's, after the broadcast) weights
        sde.rng = "torch"
        sde.set_mu(lq.to(dev))
        one = sde.reverse_sde(xT.to(dev), zs=zs.to(dev))
        sde.rng, sde.image_base = "philox", 0
        one_p = sde.reverse_sde(xT.to(dev))
        torch.save({"one": one.cpu(), "one_p": one_p.cpu()}, os.path.join(tmpdir, "single.pt"))


@pytest.mark.parametrize("precision,B", [("fp32", 5), ("bf16", 4)])
def test_native_comm_sharded_equals_single(tmp_path, precision, B):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, str(tmp_path), B, precision), nprocs=2, join=True)
    single = torch.load(tmp_path / "single.pt", weights_only=True)
    for r in range(2):
        out = torch.load(tmp_path / ("out%d.pt" % r), weights_only=True)
        assert torch.equal(out["full"], single["one"]), "rank %d: gathered sharded chain != single-GPU chain" % r
        assert torch.equal(out["full_p"], single["one_p"]), "rank %d (philox)" % r
