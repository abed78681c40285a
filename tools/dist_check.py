"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): the row-partitioned transient solve against the
single-GPU solve of the same problem.  Prints 'DIST_CHECK OK' on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/dist_check.py [cells] [steps]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    mesh, model, solver = bench.build_case(cells, rank, world)
    solver.prepare()
    infos = [solver.step(i) for i in range(1, steps + 1)]
    x, v, a = solver.owned_state()
    lp = solver.lp
    # gather the owned states on rank 0
    sizes = [int(4 * (lp.bounds[r + 1] - lp.bounds[r])) for r in range(world)]
    n_max = max(sizes)
    full = []
    for t in (x, v, a):
        pad = torch.zeros(n_max, dtype=torch.float64, device="cuda")
        pad[:t.numel()] = t
        out = [torch.zeros(n_max, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(out, pad)
        full.append(torch.cat([o[:s] for o, s in zip(out, sizes)]).cpu().numpy())
    ok = True
    if rank == 0:
        _, _, ref = bench.build_case(cells, 0, 1)
        ref.prepare()
        ref_infos = [ref.step(i) for i in range(1, steps + 1)]
        rx = (ref._x + ref.model._g_fill).cpu().numpy()
        rv, ra = ref._v.cpu().numpy(), ref._a.cpu().numpy()
        for name, got, want in (("x", full[0], rx), ("v", full[1], rv), ("a", full[2], ra)):
            idx = np.arange(want.size)
            for fld, m in (("u", idx % 4 != 3), ("theta", idx % 4 == 3)):
                err = np.linalg.norm(got[m] - want[m]) / max(np.linalg.norm(want[m]), 1e-300)
                print(f"{name}_{fld}: rel diff {err:.3e}")
                ok = ok and err < 1e-7
        print("iterations dist", [i["iterations"] for i in infos], "single", [i["iterations"] for i in ref_infos])
        print("DIST_CHECK OK" if ok else "DIST_CHECK FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
