"""BASELINE config 4 on N GPUs (run under torchrun): 64 independent 100-KF windows, window w -> rank w % N, one NCCL
all-gather of the result slots.  Checks that every rank ends up with every window's result and that those are bit-identical
to the same windows solved on one GPU without a communicator.  Writes gpurun_out/multi_gpu_check_<N>.json (rank 0)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from ms_slam_b200 import msgen, dist as msd
from ms_slam_b200.engine import Engine

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = msgen.CONFIGS["c4"]["N"]
views = [msgen.make_config("c4", 1000 + w)[0] for w in range(nwin)]
eng = Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local)
eng.comm_init(msd.broadcast_unique_id(eng, rank), rank, world)
res = eng.solve_batch(views)                       # sharded + all-gathered
ts = []
for _ in range(5):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.solve_batch(views)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
solo = Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local)
ref = solo.solve_batch(views)                      # every window on this GPU alone
same = all(np.array_equal(a.keep_bits, b.keep_bits) and np.array_equal(a.kf_cov, b.kf_cov) and np.array_equal(a.kf_slack, b.kf_slack)
           and a.objective == b.objective and a.rounds == b.rounds for a, b in zip(res, ref))
flag = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
digest = torch.tensor([int(np.bitwise_xor.reduce(np.concatenate([r.keep_bits for r in res]).astype(np.uint64))) & 0x7FFFFFFF], device="cuda")
dmax, dmin = digest.clone(), digest.clone()
dist.all_reduce(dmax, op=dist.ReduceOp.MAX); dist.all_reduce(dmin, op=dist.ReduceOp.MIN)
if rank == 0:
    out = dict(world=world, nwin=nwin, all_ranks_match_single_gpu=bool(flag.item()), ranks_agree=bool(dmax.item() == dmin.item()),
               host_call_ms_median=float(np.median(ts) * 1e3), windows_per_s=float(nwin / np.median(ts)))
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_{world}.json"), "w"))
dist.barrier()
dist.destroy_process_group()
