"""Multi-GPU check of parallel.ShardedRenderer on real hardware (run under torchrun, one rank per GPU):
every gather mode x frame dtype renders a ragged clip; every rank checks the gathered clip against its own re-render of
ALL ranks' chunks (same kernels, same chunking -> bit-exact) and rank 0 checks three frames against the oracle (1e-3).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/gpu_scripts/sharded_check.py
"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from livespeechportraits_b200 import Feature2Face_G, ShardedRenderer, partition  # noqa: E402
from livespeechportraits_b200.parallel import chunk_schedule  # noqa: E402
from oracle import f2f_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    variant, H = "normal", 256
    net = Feature2Face_G(types.SimpleNamespace(isTrain=False, size=variant, n_downsample_G=8, ngf=64, fp16=0), precision="parity")
    sd = O.make_state_dict(variant, "B")
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    n_total, chunk = 8 * world + 3, 4                       # ragged: the first ranks hold one frame more
    fm_all, cand = O.make_inputs(n_total, H, H, seed=31)
    for i in range(n_total):
        fm_all[i] = torch.roll(fm_all[i], shifts=(2 * i, 3 * i), dims=(1, 2))
    cand_d = cand[:1].to(dev)
    s, e = partition(n_total, world, rank)
    ok_all = True
    for uint8 in (False, True):
        fn = (lambda f, o: net.render_image(f, cand_d, out=o)) if uint8 else (lambda f, o: net.render(f, cand_d, out=o))
        # expected clip: every rank's chunks re-rendered locally with the same chunking
        shape = (n_total, H, H, 3) if uint8 else (n_total, 3, H, H)
        exp = torch.empty(shape, dtype=torch.uint8 if uint8 else torch.float32, device=dev)
        n_max = partition(n_total, world, 0)[1]
        for r in range(world):
            rs, re_ = partition(n_total, world, r)
            for off, ln in chunk_schedule(n_max, chunk):
                mine = max(0, min(ln, (re_ - rs) - off))
                if mine > 0:
                    fn(fm_all[rs + off:rs + off + mine].to(dev), exp[rs + off:rs + off + mine])
        torch.cuda.synchronize()
        for mode in (("ce", "nccl") if uint8 else ("ce", "mc", "nccl")):
            try:
                sr = ShardedRenderer(fn, chunk=chunk, uint8=uint8, gather=mode,
                                     render_ptr_fn=(lambda f, ptr: net.render_into_ptr(f, cand_d, ptr)) if mode == "mc" else None)
                host = torch.empty(shape, dtype=exp.dtype).pin_memory() if rank == 0 else None
                t0 = time.perf_counter()
                got = sr.render(n_total, fm_all[s:e].to(dev), host_out=host, to_host=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                same = bool(torch.equal(got, exp))
                host_ok = bool(torch.equal(host, exp.cpu())) if rank == 0 else True
                again = sr.render(n_total, fm_all[s:e].to(dev))          # reuse of the symmetric buffer, no host delivery
                torch.cuda.synchronize()
                same2 = bool(torch.equal(again, exp))
                print(f"[rank {rank}] uint8={uint8} gather={mode} (ran as {sr.gather_mode}): clip==expected {same}, host copy {host_ok}, "
                      f"second render {same2}, first call {dt * 1e3:.1f} ms", flush=True)
                ok_all = ok_all and same and host_ok and same2
            except Exception as exc:      # noqa: BLE001
                print(f"[rank {rank}] uint8={uint8} gather={mode}: FAILED {type(exc).__name__}: {exc}", flush=True)
                if not (mode == "mc" and "unavailable" in str(exc)):       # no NVLink multicast on this node: reported, not a failure
                    ok_all = False
        if rank == 0 and not uint8:
            for i in (0, n_total // 2, n_total - 1):
                x = torch.cat([fm_all[i:i + 1], cand[:1]], 1)
                err = (exp[i:i + 1].cpu() - O.generator_forward(sd, x, variant)).abs().max().item()
                print(f"[rank 0] frame {i}: max|cuda - oracle| = {err:.3g}", flush=True)
                ok_all = ok_all and err <= 1e-3
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
