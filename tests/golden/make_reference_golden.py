"""Generates tests/golden/reference_fmax_32.npz: outputs of the REFERENCE's own compute_fmax()
(src/fmax.c, fmax-pfft.c, LPT.c, collapse_times.c compiled verbatim into oracle/_ref, see
oracle/Makefile) for a seeded 32^3 box.  Needs /root/reference; run from the repo root:

    make -C oracle && OMP_NUM_THREADS=8 python tests/golden/make_reference_golden.py

The input field comes from the oracle's GenIC restatement (the reference's GenIC needs GSL's
generators and is pinned separately by the sigma(R)/FmaxPDF fixtures of hmf_validation/).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import pinocchio_oracle as po  # noqa: E402
from oracle.reference_runner import HMF_RADII, ReferenceRun  # noqa: E402
from pinocchio_b200.cosmology import Cosmology  # noqa: E402

N, SEED = 32, 486604


def main():
    cosmo = Cosmology(pk_norm_override=2.03146e7)
    box = N / 0.7
    kd = po.genic(N, box, SEED, cosmo.PowerSpectrum)
    growth = np.array([cosmo.GrowingMode(0.0), cosmo.GrowingMode_2LPT(0.0), cosmo.GrowingMode_3LPT_1(0.0),
                       cosmo.GrowingMode_3LPT_2(0.0)])
    run = ReferenceRun(N, box, HMF_RADII, growth, cosmo.sp_invgrow.x, cosmo.sp_invgrow.y, threads=8)
    run.set_kdensity(kd)
    _, tv = run.compute_fmax()
    prod = run.products(po.PRODUCT_DTYPE_3LPT)
    pdf = run.fmax_pdf_file()
    out = Path(__file__).resolve().parent / "reference_fmax_32.npz"
    np.savez_compressed(out, N=N, seed=SEED, box=box, radii=np.array(HMF_RADII), growth=growth,
                        invgrow_x=np.asarray(cosmo.sp_invgrow.x), invgrow_y=np.asarray(cosmo.sp_invgrow.y),
                        kdensity=kd, products=prod.view(np.uint8), true_variance=tv, fmax_pdf_file=pdf,
                        kvector_2LPT=run.kvector(0), kvector_3LPT_1=run.kvector(1), kvector_3LPT_2=run.kvector(2))
    print(out, out.stat().st_size, "bytes; collapsed cells:", int((prod["Fmax"] >= 1.0).sum()))


if __name__ == "__main__":
    main()
