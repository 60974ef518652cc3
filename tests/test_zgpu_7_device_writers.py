"""On-disk formats written straight from the device SoA: pinb200_write_products (the
DumpProducts/Task.<rank> records of src/fmax.c:372-426) and pinb200_write_block (the FMAX / RMAX / ZEL / 2LPT /
31PT / 32PT payloads of the timeless snapshot, src/write_snapshot.c:695-860), through pinned staging without a
host products[] array.  Byte-exact against what the reference's own loops would write from the downloaded
fields; then the drop-in PROGRAM's DumpProducts directory against the reference program's.  Needs a B200: -m gpu.
"""
import os
from pathlib import Path

import numpy as np
import pytest

from test_reference_full import REF_X, run_program

pytestmark = pytest.mark.gpu
HMF_RADII = [20.635922, 13.996056, 9.026099, 5.465945, 3.058354, 1.548258, 0.689079, 0.258729, 0.0]
B200_X = REF_X.parent / "pinocchio_b200.x"


def make(N, lpt_order=3):
    from pinocchio_b200.cosmology import Cosmology, SmoothingLadder
    from pinocchio_b200.engine import Pinocchio, RunConfig
    cfg = RunConfig(GridSize=N, BoxSize_htrue=N / 0.7, lpt_order=lpt_order)
    return Pinocchio(cfg, Cosmology(pk_norm_override=2.03146e7), smoothing=SmoothingLadder(np.array(HMF_RADII), np.zeros(9)))


def written(tmp_path, name, fn):
    path = tmp_path / name
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        fn(fd)
    finally:
        os.close(fd)
    return path.read_bytes()


@pytest.mark.parametrize("N", [32, 128])
def test_blocks_and_products_from_device(N, tmp_path):
    p = make(N)
    p.GenIC_large()
    p.compute_fmax()
    ncell = N ** 3
    # blocks: what initialize_FMAX / _RMAX / _ZEL / ... put into block.data (one float, one int, or AuxStruct {float axis[3]})
    assert written(tmp_path, "fmax", lambda fd: p.write_block(fd, "FMAX")) == p.field("Fmax").tobytes()
    assert written(tmp_path, "rmax", lambda fd: p.write_block(fd, "RMAX")) == p.field("Rmax").tobytes()
    for block, name in (("ZEL ", "Vel"), ("2LPT", "Vel_2LPT"), ("31PT", "Vel_3LPT_1"), ("32PT", "Vel_3LPT_2")):
        want = np.stack([p.field(name, a).ravel() for a in range(3)], axis=1)
        assert written(tmp_path, "blk", lambda fd: p.write_block(fd, block)) == want.tobytes(), block
    # a sub-range (a task of a multi-file snapshot writes its own particles only)
    b0, n = ncell // 3 + 5, ncell // 2 + 1
    want = np.stack([p.field("Vel_2LPT", a).ravel() for a in range(3)], axis=1)[b0:b0 + n]
    assert written(tmp_path, "blk", lambda fd: p.write_block(fd, "2LPT", b0, n)) == want.tobytes()
    # the dump: fwrite(products, sizeof(product_data), total_local_size, file)
    prod = p.products()
    assert written(tmp_path, "task", lambda fd: p.write_products(fd)) == prod.tobytes()
    assert written(tmp_path, "task", lambda fd: p.write_products(fd, b0, n)) == prod[b0:b0 + n].tobytes()
    # appended to what the caller has already written (block header, earlier tasks)
    assert written(tmp_path, "app", lambda fd: (os.write(fd, b"HEAD"), p.write_block(fd, "FMAX", 0, 7))) == b"HEAD" + p.field("Fmax").ravel()[:7].tobytes()
    p.close()


def test_writers_many_chunks_and_errors(tmp_path):
    """256^3: eight staging rounds per record stream; error paths"""
    from pinocchio_b200.engine import PinocchioError
    N = 256
    p = make(N, lpt_order=2)
    p.GenIC_large()
    p.compute_fmax()
    want = np.stack([p.field("Vel", a).ravel() for a in range(3)], axis=1)
    assert written(tmp_path, "zel", lambda fd: p.write_block(fd, "ZEL ")) == want.tobytes()
    from pinocchio_b200.engine import PRODUCT_DTYPE_3LPT
    prod = p.products()                       # 56-byte records; the 3LPT members stay zero with lpt_order = 2
    assert not prod["Vel_3LPT_1"].any()
    assert written(tmp_path, "task", lambda fd: p.write_products(fd)) == prod.tobytes()
    with pytest.raises(PinocchioError):
        written(tmp_path, "x", lambda fd: p.write_block(fd, "31PT"))          # field of a higher LPT order: not resident
    with pytest.raises(PinocchioError):
        written(tmp_path, "x", lambda fd: p.write_block(fd, "FMAX", N ** 3 - 3, 4))    # range outside the slab
    rd = os.open(tmp_path / "ro", os.O_RDONLY | os.O_CREAT, 0o644)
    with pytest.raises(PinocchioError):
        p.write_block(rd, "FMAX")                                               # write(2) fails: reported, not ignored
    os.close(rd)
    p.close()


@pytest.mark.skipif(not (REF_X.exists() and B200_X.exists()), reason="oracle/_ref/pinocchio_{ref,b200}.x not built (make -C oracle all)")
def test_dropin_program_dump_against_reference_program(tmp_path):
    """`DumpProducts` through the linked drop-in (shim dump_products -> pinb200_write_products: Task.0 written from
    the device while products[] on the host holds the compact hand-off only) against the reference program's dump:
    same summary, same TrueVariance to 1e-12, records equal to float rounding (1e-6, a few ill-conditioned cells)."""
    from oracle import pinocchio_oracle as po

    def run(exe, d):
        d.mkdir()
        for name in ("parameter_file", "outputs"):
            (d / name).write_bytes((Path(__file__).parent / "golden" / "hmf_validation" / name).read_bytes())
        with open(d / "parameter_file", "a") as f:
            f.write("\nDumpProducts\n")
        import subprocess
        r = subprocess.run([str(exe), "parameter_file"], cwd=d, capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, OMP_NUM_THREADS="16"))
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        return r.stdout

    la = run(B200_X, tmp_path / "b200")
    run(REF_X, tmp_path / "ref")
    assert "compact hand-off" in la
    da, db = tmp_path / "b200" / "DumpProducts", tmp_path / "ref" / "DumpProducts"
    assert (da / "summary").read_text() == (db / "summary").read_text()
    tva, tvb = np.fromfile(da / "TrueVariance"), np.fromfile(db / "TrueVariance")
    assert tva.size == tvb.size == 9 and np.abs(tva / tvb - 1).max() < 1e-12
    a = np.fromfile(da / "Task.0", dtype=po.PRODUCT_DTYPE_3LPT)
    b = np.fromfile(db / "Task.0", dtype=po.PRODUCT_DTYPE_3LPT)
    assert a.size == b.size == 128 ** 3
    assert (a["Rmax"] != b["Rmax"]).mean() < 1e-3
    dF = np.abs(a["Fmax"].astype(np.float64) - b["Fmax"]) / np.maximum(1.0, np.abs(b["Fmax"]))
    assert (dF > 1e-6).mean() < 2e-4 and (a["Fmax"] == b["Fmax"]).mean() > 0.99
    for name in ("Vel", "Vel_2LPT", "Vel_3LPT_1", "Vel_3LPT_2"):
        scale = np.abs(b[name]).max()
        assert np.abs(a[name].astype(np.float64) - b[name]).max() <= 1e-6 * scale, name
        assert (a[name] == b[name]).mean() > 0.99, name
