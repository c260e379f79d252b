"""One rank of the 2-GPU test of the library's own NCCL exchange (tests/test_zx_gpu_comm.py launches two of these).
No torch.distributed: the 128-byte NCCL id travels through a file, as an MPI host would broadcast it."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, out_dir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import torch   # device memory for the device-resident step only — imported BEFORE the library binds libnccl.so.2, so that both use torch's copy
    from sdfibm_b200 import cases
    from sdfibm_b200.context import Context

    n, n_solids, n_side = 48, 60, 4
    if rank == 1:
        os.environ["SDFIBM_HEAVY_CAP0"] = "64"   # this rank's first step overflows its queue and runs again: the flag protocol
    case = cases.case_c5_block(rank, world, n=n, n_solids=n_solids, n_side=n_side)
    ctx = Context(rank)
    ctx.set_mesh(case["mesh"], False)
    ctx.set_shapes(case["shapes"])
    id_file = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        uid = ctx.comm_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise SystemExit("no NCCL id")
            time.sleep(0.05)
        uid = open(id_file, "rb").read()
    ctx.comm_init(uid, rank, world)
    # 0. the device-resident entry (one CUDA-graph launch per step): the SPLIT step — the all-reduce of the sums on the library's
    #    communication stream alongside the certificate pass, the ranks' retry flags in a second, 8-byte all-reduce.  Rank 1's queue
    #    overflows in this very step, so the flag protocol and the second reduction run in split mode.
    dev = torch.device("cuda", rank)
    nC, nS = case["mesh"].n_cells, len(case["solids"])
    dU = torch.from_numpy(case["U"]).to(dev)
    f = [torch.zeros(k, dtype=torch.float64, device=dev) for k in (nC, 3 * nC, nC, nC, 6 * nS)]
    from sdfibm_b200 import capi
    solids_pinned = capi.pinned_like(np.ascontiguousarray(case["solids"], dtype=capi.SOLID_DTYPE))
    dev_ft = []
    for _ in range(3):     # the first step carries the retry; the next two replay the captured graph
        ctx.interact_device(solids_pinned, dU.data_ptr(), case["dt"], case["rhof"], *[x.data_ptr() for x in f])
        dev_ft.append(f[4].cpu().numpy().reshape(nS, 6).copy())
    dev_As = f[0].cpu().numpy().copy()
    # 1. the default: slice upload + all-gather of the replicated solids, force/torque summed over the ranks
    total = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    lists = ctx.candidate_lists()
    # 2. the same step with the exchange switched off: this rank's partial sums
    ctx.comm_options(auto_reduce=False, gather_solids=False)
    part = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    # 3. once more with the exchange on (step 1 included rank 1's local retry; this one does not)
    ctx.comm_options(auto_reduce=True, gather_solids=True)
    again = ctx.interact(case["solids"], case["U"], case["dt"], case["rhof"])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), FT=total["FT"], FT_partial=part["FT"], FT_again=again["FT"], As=total["As"],
             As_partial=part["As"], Ct=total["Ct"], off=lists[0], cells=lists[1], comm_ms=ctx.comm_last_ms(),
             FT_dev0=dev_ft[0], FT_dev1=dev_ft[1], FT_dev2=dev_ft[2], As_dev=dev_As)
    ctx.close()


if __name__ == "__main__":
    main()
