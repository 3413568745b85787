"""torchrun target: frame-sharded U-Net forward over NCCL vs the single-GPU forward and the CPU oracle (rank 0 prints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from oracle import lvdm_oracle as O, synth
from viewcrafter_b200 import parallel
from viewcrafter_b200.configs import UNET_PARAMS
from viewcrafter_b200.unet import UNetModel

m = UNetModel(**dict(UNET_PARAMS, model_channels=64))
sd = synth.synth_state_dict(synth.module_shapes(m), 5)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
g = torch.Generator().manual_seed(6)
B, T, H, W = 2, 5, 16, 16
x, ctx = torch.randn(B, 8, T, H, W, generator=g), torch.randn(B, 333, 1024, generator=g)
t = torch.tensor([499, 19])
y_single = m(x.cuda(), t.cuda(), context=ctx.cuda())
comm = parallel.shard_model(m, dist, rank, world)
y_sharded = m(x.cuda(), t.cuda(), context=ctx.cuda())
torch.cuda.synchronize()
d_single = float((y_sharded - y_single).abs().max())
ok = d_single < 0.02
if rank == 0:
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, ctx, None, default_fs=10)
    d_ref = float((y_sharded.cpu() - ref).abs().max())
    ok = ok and d_ref < 0.02
    print(f"world {world}: |sharded - single| {d_single:.4g}, |sharded - oracle| {d_ref:.4g}, all-to-all bytes sent by rank 0: {comm.bytes_moved}")
    if ok:
        print("PARALLEL_CHECK_OK")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
