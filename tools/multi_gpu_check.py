"""Run under torchrun (one rank per GPU): N-agent ADMM with one agent per GPU, consensus exchange through the
library's in-stream ncclAllReduce; rank 0 checks every agent's solution against the oracle's in-process ADMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from d2slam_b200 import abi, synth
from d2slam_b200.solver import Solver, comm_unique_id

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
steps, iters = 2, 6
sw = synth.make_swarm(seed=5, n_agents=world, n_landmarks=120, shared_per_pair=20)
cfg = dict(consensus_max_steps=steps, max_num_iterations=iters)
s = Solver(device=lr, **cfg)
sw[rank].load(s, 0)
s.finalize()
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.tensor(list(comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
s.comm_init(bytes(uid.cpu().tolist()), rank, world)
rep = s.solve_fixed(iters)[0]
pose = torch.from_numpy(s.get_blocks(0, abi.POSE, sw[rank]["frame_ids"])).cuda()
lm = torch.from_numpy(s.get_blocks(0, abi.LANDMARK, sw[rank]["lm_ids"])[:, 0].copy()).cuda()
poses = [torch.zeros_like(pose) for _ in range(world)]; lms = [torch.zeros_like(lm) for _ in range(world)]
dist.all_gather(poses, pose); dist.all_gather(lms, lm)
ok = True
if rank == 0:
    from oracle import orc
    ags = []
    for p in sw:
        a = orc.Oracle(**cfg); p.load(a); ags.append(a)
    orc.admm_solve(ags, fixed_mode=True)
    for a in range(world):
        po = ags[a].get_blocks(abi.POSE, sw[a]["frame_ids"])
        dp, dr = synth.pose_errors(poses[a].cpu().numpy(), po)
        lo = ags[a].get_blocks(abi.LANDMARK, sw[a]["lm_ids"])[:, 0]
        dl = np.abs(lms[a].cpu().numpy() / lo - 1).max()
        print(f"agent {a}: pose diff {dp:.3e} m {dr:.3e} rad, landmark rel {dl:.3e}")
        ok = ok and dp < 1e-6 and dr < 1e-6 and dl < 1e-5
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", f"world={world} t={rep.total_time*1e3:.3f} ms")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
