"""Multi-GPU parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py [N]
Every rank owns one x-slab; rank 0 gathers the slabs and compares the global fields with the
oracle (same tolerances as tests/test_gpu_parity.py).  Exit code 0 = parity green.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]


def gather(a, axis):
    """all ranks -> rank 0 concatenation along `axis` (through the default process group)."""
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy() for o in outs], axis=axis)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import pinocchio_oracle as po
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, RunConfig

    cosmo = Cosmology(pk_norm_override=2.03146e7)
    cfg = RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=3)
    pin = Pinocchio(cfg, cosmo, device=local, smoothing=SmoothingLadder(np.array(HMF_RADII), np.zeros(9)),
                    rank=rank, nranks=world)
    pin.GenIC_large()
    kd = gather(pin.read_kdensity().view(np.float64), axis=1).view(np.complex128)     # K layout: split along y
    pin.compute_fmax()
    Fmax = gather(pin.field("Fmax"), 0)
    Rmax = gather(pin.field("Rmax"), 0)
    vel = {n: [gather(pin.field(n, a), 0) for a in range(3)] for n in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2")}
    kvec = [gather(pin.read_kvector(w).view(np.float64), axis=1).view(np.complex128) for w in range(3)]
    pdf = pin.Fmax_PDF()
    tv = pin.TrueVariance
    # forward_transform / reverse_transform on the slabs (collective, src/fmax-pfft.c:191-228)
    rng = np.random.default_rng(5)
    rfull = rng.standard_normal((N, N, N))
    lx = N // world
    ck = gather(pin.forward_transform(rfull[rank * lx:(rank + 1) * lx]).view(np.float64), axis=1).view(np.complex128)
    cfull = rng.standard_normal((N, N, N // 2 + 1)) + 1j * rng.standard_normal((N, N, N // 2 + 1))     # not Hermitian
    rback = gather(pin.reverse_transform(cfull[:, rank * lx:(rank + 1) * lx]), 0)
    ok = True
    if rank == 0:
        e1 = np.abs(ck - np.fft.rfftn(rfull)).max() / np.abs(ck).max()
        want = np.fft.irfftn(cfull, s=(N, N, N), axes=(0, 1, 2))
        e2 = np.abs(rback - want).max() / np.abs(want).max()
        print(f"[multi {world} GPUs, {N}^3] slab r2c rel err {e1:.2e}, c2r rel err {e2:.2e}")
        ok &= e1 < 1e-14 and e2 < 1e-14
    if rank == 0:
        ref_kd = po.genic(N, N / 0.7, 486604, cosmo.PowerSpectrum)
        e = np.abs(kd - ref_kd).max() / np.abs(ref_kd).max()
        print(f"[multi {world} GPUs, {N}^3] kdensity rel err {e:.2e}")
        ok &= e < 1e-13
        g = tuple(pin.growth_rates(0.0))
        ref = po.compute_fmax(kd, HMF_RADII, 1.0 / 0.7, cosmo.InverseGrowingMode, growth=g, keep=True)
        e = np.abs(tv / ref["TrueVariance"] - 1).max()
        print(f"  TrueVariance rel err {e:.2e}")
        ok &= e < 1e-12
        good = ~ref["unstable"]
        dF = np.abs(Fmax.astype(np.float64) - ref["Fmax"])
        okF = (dF[good] <= 1e-6 * np.maximum(1.0, np.abs(ref["Fmax"][good]))).all()
        top2 = np.sort(np.stack(ref["F"]), axis=0)[-2:]
        ties = np.abs(top2[1] - top2[0]) <= 1e-6 * np.maximum(1.0, np.abs(top2[1]))
        badR = ((Rmax != ref["Rmax"]) & ~ties & good).sum()
        print(f"  Fmax ok {okF}, max dF {dF[good].max():.2e}, ill-conditioned cells {int((~good).sum())}, Rmax mismatches {badR}")
        ok &= bool(okF) and badR == 0
        pe = np.abs(pdf.astype(np.int64) - po.fmax_pdf(ref["Fmax"]).astype(np.int64)).max()
        print(f"  FmaxPDF max bin diff {pe}, total {pdf.sum()}")
        ok &= pe <= 2 + 2 * int((~good).sum()) and pdf.sum() == N ** 3
        for w, name in enumerate(("kvector_2LPT", "kvector_3LPT_1", "kvector_3LPT_2")):
            e = np.abs(kvec[w] - ref[name]).max() / np.abs(ref[name]).max()
            print(f"  {name} rel err {e:.2e}")
            ok &= e < 1e-11
        for n in vel:
            for a in range(3):
                e = np.abs(vel[n][a].astype(np.float64) - ref[n][a]).max() / np.abs(ref[n][a]).max()
                ok &= e <= 1e-6
            print(f"  {n} ok (<= 1e-6 of max)")
        print("MULTI-GPU PARITY", "GREEN" if ok else "RED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    pin.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
