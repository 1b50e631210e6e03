"""Run the UNMODIFIED reference CPU pair styles (oracle/_ref/ref_*) on a Config.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The binaries are built by oracle/Makefile
from /root/reference sources against oracle/shim; on the GPU box the prebuilt files are used.
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = {
    "annp_fe": os.path.join(HERE, "_ref", "ref_annp_fe"),
    "annp_ni": os.path.join(HERE, "_ref", "ref_annp_ni"),
    "anna_adp": os.path.join(HERE, "_ref", "ref_anna_adp"),
    # NOT an oracle: our own LAMMPS-facing pair style (meng_zhang_b200/lammps/pair_annp_b200.cpp) linked against
    # libannp_b200.so and driven by the same shim + driver, so it is tested exactly like the reference style
    "plugin_annp_b200": os.path.join(HERE, "_ref", "plugin_annp_b200"),
    "plugin_annp_ni_b200": os.path.join(HERE, "_ref", "plugin_annp_ni_b200"),
    "plugin_anna_adp_b200": os.path.join(HERE, "_ref", "plugin_anna_adp_b200"),
}


def available(kind: str) -> bool:
    if os.path.isfile(BIN[kind]) and not os.access(BIN[kind], os.X_OK):
        try:                       # the snapshot that ships the binaries to the GPU box may drop the mode bits
            os.chmod(BIN[kind], 0o755)
        except OSError:
            pass
    return os.path.isfile(BIN[kind]) and os.access(BIN[kind], os.X_OK)


def write_input(path, cfg, eflag=3, vflag=2, ncalls=1):
    with open(path, "wb") as fp:
        fp.write(struct.pack("<8i", cfg.nlocal, cfg.nghost, int(cfg.type.max()), len(cfg.ilist),
                             eflag, vflag, ncalls, 0))
        fp.write(np.ascontiguousarray(cfg.x, dtype="<f8").tobytes())
        fp.write(np.ascontiguousarray(cfg.type, dtype="<i4").tobytes())
        fp.write(np.ascontiguousarray(cfg.ilist, dtype="<i4").tobytes())
        fp.write(np.ascontiguousarray(cfg.numneigh, dtype="<i4").tobytes())
        fp.write(np.ascontiguousarray(cfg.neigh, dtype="<i4").tobytes())
        fp.write(np.ascontiguousarray(cfg.ghost_owner, dtype="<i4").tobytes())


def read_output(path):
    with open(path, "rb") as fp:
        nall, has_e, has_v, ncalls = struct.unpack("<4i", fp.read(16))
        eng = struct.unpack("<d", fp.read(8))[0]
        virial = np.frombuffer(fp.read(48), dtype="<f8").copy()
        secs = struct.unpack("<d", fp.read(8))[0]
        f = np.frombuffer(fp.read(nall * 24), dtype="<f8").reshape(nall, 3).copy()
        eatom = np.frombuffer(fp.read(nall * 8), dtype="<f8").copy() if has_e else None
        vatom = np.frombuffer(fp.read(nall * 48), dtype="<f8").reshape(nall, 6).copy() if has_v else None
        per_call = np.frombuffer(fp.read(ncalls * 8), dtype="<f8").copy()
    return {"eng_vdwl": eng, "virial": virial, "f": f, "eatom": eatom, "vatom": vatom,
            "seconds": secs, "ncalls": ncalls, "per_call_seconds": per_call}


def run_reference(kind, cfg, potential_file, elements, eflag=3, vflag=2, ncalls=1, timeout=3600, newton=None, env_extra=None):
    """One `Pair::compute(eflag, vflag)` of the reference style `kind` on cfg.

    Returns eng_vdwl, virial[6], f[nall,3] (ghost forces NOT folded), eatom, vatom, seconds."""
    if not available(kind):
        raise FileNotFoundError(f"{BIN[kind]} not built: run `make -C oracle`")
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        write_input(fin, cfg, eflag, vflag, ncalls)
        cmd = [BIN[kind], fin, fout, potential_file] + list(elements)
        env = dict(os.environ)
        if newton is not None:
            env["ANNP_DRIVER_NEWTON"] = str(int(newton))      # the deck's `newton on|off`
        env.update(env_extra or {})
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        if p.returncode != 0:
            raise RuntimeError(f"reference driver failed ({p.returncode}): {p.stderr[-2000:]}")
        out = read_output(fout)
    out["stdout"] = p.stdout
    return out
