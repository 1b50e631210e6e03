"""Run a LAMMPS input deck of the kind the reference ships (annp-gpu-lammps/fe_v2 `performance test.zip`: in.st_test)
on the device-resident driver, without LAMMPS:

    python -m meng_zhang_b200.deck in.st_test            [under torch.distributed.run for one rank per GPU]

Only the commands those decks use are understood - enough to take the reference's own deck verbatim:
    echo, processors, package, newton, units metal, atom_style atomic, boundary, timestep, neighbor, neigh_modify,
    variable (equal / string), read_data (atomic), pair_style annp[/gpu] | anna_adp[/gpu], pair_coeff, mass, min_style cg,
    minimize, reset_timestep, thermo, thermo_style custom, velocity all create, fix nve | nvt | npt, unfix, run
`dump` / `dump_modify` are accepted and ignored (a note is printed).  Anything else is an error, not a silent skip.
What LAMMPS itself contributes to such a run is restated in lammps_compat.py (velocity generator, shrink-wrapped box),
md.py (minimiser, neighbour trigger) and csrc/annp_nh.cu (fix nvt / npt); the forces are libannp_b200.so.
Thermo lines are printed in LAMMPS' layout so that a log can be diffed against a LAMMPS log.
"""
from __future__ import annotations

import math
import os
import re
import sys

import numpy as np

THERMO_KEYS = ("step", "temp", "pe", "ke", "etotal", "lx", "ly", "lz", "press", "vol", "pxx", "pyy", "pzz")
THERMO_HEAD = {"step": "Step", "temp": "Temp", "pe": "PotEng", "ke": "KinEng", "etotal": "TotEng", "lx": "Lx", "ly": "Ly", "lz": "Lz",
               "press": "Press", "vol": "Volume", "pxx": "Pxx", "pyy": "Pyy", "pzz": "Pzz"}


class DeckError(RuntimeError):
    pass


def read_data_atomic(path):
    """`read_data` for atom_style atomic: header counts, box bounds, `Atoms` section (id type x y z).
    Returns x[n,3] in atom-ID order, type[n], box[3,2], ntypes."""
    with open(path) as fp:
        lines = fp.read().splitlines()
    natoms = ntypes = None
    box = np.zeros((3, 2))
    i = 1                                           # first line is a title
    while i < len(lines):
        t = lines[i].split("#")[0].split()
        if len(t) >= 2 and t[1] == "atoms":
            natoms = int(t[0])
        elif len(t) >= 3 and t[1] == "atom" and t[2] == "types":
            ntypes = int(t[0])
        elif len(t) >= 4 and t[2] in ("xlo", "ylo", "zlo"):
            box["xyz".index(t[2][0])] = [float(t[0]), float(t[1])]
        elif t and t[0] == "Atoms":
            break
        i += 1
    if natoms is None or ntypes is None:
        raise DeckError(f"{path}: not a LAMMPS data file (atoms / atom types missing)")
    rows = []
    i += 1
    while len(rows) < natoms and i < len(lines):
        t = lines[i].split()
        if len(t) >= 5:
            rows.append(t[:5])
        i += 1
    if len(rows) != natoms:
        raise DeckError(f"{path}: expected {natoms} atoms, found {len(rows)}")
    a = np.array(rows, dtype=np.float64)
    order = np.argsort(a[:, 0].astype(np.int64), kind="stable")
    return a[order, 2:5], a[order, 1].astype(np.int32), box, ntypes


class Deck:
    def __init__(self, out=sys.stdout, device_index=None):
        self.out = out
        self.vars = {}
        self.dt, self.skin = 0.001, 2.0
        self.every, self.delay, self.check = 1, 0, True           # LAMMPS defaults (neigh_modify)
        self.boundary = ["p", "p", "p"]
        self.x = self.types = self.box = None
        self.masses = {}
        self.pair = self.md = None
        self.pair_style = None
        self.thermo_every = 0
        self.thermo_cols = ["step", "temp", "pe", "ke", "etotal", "press"]     # thermo_style one (ke, not in one, is harmless)
        self.fix = None
        self.step = 0
        self.rank, self.world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        self.device_index = device_index
        self.cwd = "."
        self.rows = []                                              # thermo rows of the last run (for callers / tests)

    # ------------------------------------------------------------------ text level
    def say(self, s):
        if self.rank == 0:
            print(s, file=self.out, flush=True)

    def substitute(self, line):
        def immediate(m):
            return repr(self.evaluate(m.group(1)))
        line = re.sub(r"\$\(([^)]*)\)", immediate, line)
        line = re.sub(r"\$\{(\w+)\}", lambda m: self._var(m.group(1)), line)
        return re.sub(r"\$(\w)", lambda m: self._var(m.group(1)), line)

    def _var(self, name):
        if name not in self.vars:
            raise DeckError(f"Substitution for illegal variable {name}")
        return str(self.vars[name])

    def evaluate(self, expr):
        """`equal`-style formulas of the decks: arithmetic over numbers, variables (v_name), the keyword dt and the functions
        sqrt / exp / ln.  The formula is parsed into a syntax tree and only numbers, + - * / ^, unary minus, the known names
        and calls of the three functions are evaluated; anything else (attributes, subscripts, other calls) is an error,
        and exponents are bounded so a formula cannot stall the run."""
        import ast
        import operator
        expr = re.sub(r"v_(\w+)", lambda m: self._var(m.group(1)), expr)
        names = {"dt": self.dt, "PI": math.pi}
        funcs = {"sqrt": math.sqrt, "exp": math.exp, "ln": math.log}
        binops = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv}
        bad = DeckError(f"Invalid syntax in variable formula: {expr}")

        def ev(node):
            if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
                return float(node.value)
            if isinstance(node, ast.Name) and node.id in names:
                return float(names[node.id])
            if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
                v = ev(node.operand)
                return -v if isinstance(node.op, ast.USub) else v
            if isinstance(node, ast.BinOp) and type(node.op) in binops:
                return binops[type(node.op)](ev(node.left), ev(node.right))
            if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Pow):
                base, power = ev(node.left), ev(node.right)
                if abs(power) > 1024.0:
                    raise bad
                return math.pow(base, power)
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in funcs and len(node.args) == 1 and not node.keywords:
                return funcs[node.func.id](ev(node.args[0]))
            raise bad

        try:
            tree = ast.parse(expr.strip().replace("^", "**"), mode="eval")
            return float(ev(tree.body))
        except (SyntaxError, ValueError, ZeroDivisionError, OverflowError, RecursionError):
            raise bad

    def run_file(self, path):
        self.cwd = os.path.dirname(os.path.abspath(path))
        with open(path) as fp:
            text = fp.read()
        self.run_text(text)

    def run_text(self, text):
        pending = ""
        for raw in text.splitlines():
            line = raw.split("#")[0].rstrip() if not raw.lstrip().startswith("#") else ""
            if line.endswith("&"):
                pending += line[:-1] + " "
                continue
            line = (pending + line).strip()
            pending = ""
            if not line:
                continue
            words = self.substitute(line).split()
            self.command(words[0], words[1:])

    # ------------------------------------------------------------------ commands
    def command(self, cmd, a):
        fn = getattr(self, "cmd_" + cmd, None)
        if fn is None:
            raise DeckError(f"Unknown command: {cmd} (meng_zhang_b200.deck understands the reference decks' subset only)")
        fn(a)

    def cmd_echo(self, a): pass

    def cmd_processors(self, a):
        want = [int(v) if v != "*" else 0 for v in a[:3]]
        self.grid_request = want

    def cmd_package(self, a): pass            # `package gpu N neigh no`: one rank per GPU here, nothing to configure

    def cmd_newton(self, a):
        if a[0] not in ("on", "off"):
            raise DeckError("Illegal newton command")
        self.newton = a[0]                      # checked against the pair style in pair_coeff

    def cmd_units(self, a):
        if a[0] != "metal":
            raise DeckError("only `units metal` is supported")

    def cmd_atom_style(self, a):
        if a[0] != "atomic":
            raise DeckError("only `atom_style atomic` is supported")

    def cmd_boundary(self, a):
        for b in a[:3]:
            if b not in ("p", "m", "s", "f"):
                raise DeckError(f"boundary {b}: only p, m, s, f on both faces are supported")
        self.boundary = list(a[:3])

    def cmd_timestep(self, a): self.dt = float(a[0])

    def cmd_neighbor(self, a): self.skin = float(a[0])

    def cmd_neigh_modify(self, a):
        for k, v in zip(a[::2], a[1::2]):
            if k == "every": self.every = int(v)
            elif k == "delay": self.delay = int(v)
            elif k == "check": self.check = (v == "yes")
            else: raise DeckError(f"neigh_modify {k} is not supported")

    def cmd_variable(self, a):
        name, style = a[0], a[1]
        if style == "equal":
            self.vars[name] = self.evaluate(" ".join(a[2:]))
            if float(self.vars[name]).is_integer():
                self.vars[name] = int(self.vars[name])
        elif style in ("string", "index"):
            self.vars[name] = a[2]
        else:
            raise DeckError(f"variable style {style} is not supported")

    def cmd_read_data(self, a):
        path = a[0] if os.path.isabs(a[0]) else os.path.join(self.cwd, a[0])
        self.x, self.types, self.box, self.ntypes = read_data_atomic(path)
        self.say(f"  orthogonal box = ({self.box[0, 0]:g} {self.box[1, 0]:g} {self.box[2, 0]:g}) to ({self.box[0, 1]:g} {self.box[1, 1]:g} {self.box[2, 1]:g})")
        self.say(f"  {len(self.x)} atoms")

    def cmd_pair_style(self, a):
        name = a[0]
        if name.split("/")[0] not in ("annp", "anna_adp"):
            raise DeckError(f"pair_style {name}: this library serves annp[/gpu] and anna_adp[/gpu]")
        if len(a) != 1:
            raise DeckError("Illegal pair_style command")
        self.pair_style = name.split("/")[0]

    def cmd_pair_coeff(self, a):
        from .pair import PairANNPGPU
        from .pair_anna import PairANNAADPGPU
        cls = PairANNPGPU if self.pair_style == "annp" else PairANNAADPGPU
        if getattr(self, "newton", "on") == "off" and self.pair_style == "annp":
            raise DeckError("Pair style annp/gpu requires newton pair on")        # fe_v2/src/pair_annp_gpu.cpp:139-140
        # anna_adp: the reference's GPU decks say `newton off`, its CPU decks `newton on`; the forces are the same
        self.pair = cls(ntypes=self.ntypes, device=-1 if self.device_index is None else self.device_index, skin=self.skin)
        self.pair.settings([])
        args = list(a)
        args[2] = args[2] if os.path.isabs(args[2]) else os.path.join(self.cwd, args[2])
        self.pair.coeff(args)
        self.pair.init_style()

    def cmd_mass(self, a): self.masses[int(a[0]) if a[0] != "*" else 0] = float(a[1])

    def cmd_min_style(self, a):
        if a[0] != "cg":
            raise DeckError("only `min_style cg` is implemented")

    def cmd_reset_timestep(self, a): self.step = int(a[0])

    def cmd_thermo(self, a): self.thermo_every = int(a[0])

    def cmd_thermo_style(self, a):
        if a[0] == "one":
            return
        if a[0] != "custom":
            raise DeckError("thermo_style: only `one` and `custom` are supported")
        for k in a[1:]:
            if k not in THERMO_KEYS:
                raise DeckError(f"thermo_style custom keyword {k} is not supported")
        self.thermo_cols = list(a[1:])

    def cmd_dump(self, a): self.say(f"(meng_zhang_b200.deck: `dump {' '.join(a)}` accepted and ignored)")

    def cmd_dump_modify(self, a): pass

    def cmd_velocity(self, a):
        import torch
        from .lammps_compat import velocity_create
        if a[0] != "all" or a[1] != "create":
            raise DeckError("velocity: only `velocity all create T seed` (default options) is supported")
        if len(a) > 4:
            raise DeckError("velocity create options beyond the defaults are not supported")
        md = self._md()
        v = velocity_create(len(self.x), self._mass(), float(a[2]), int(a[3]))
        md.v = torch.as_tensor(v[md.gid.cpu().numpy()], dtype=torch.float64, device=md.dev)

    def cmd_fix(self, a):
        if a[1] != "all":
            raise DeckError("fix: only group `all` is supported")
        style, args = a[2], a[3:]
        if style == "nve":
            self.fix = (a[0], "nve", {})
            return
        if style not in ("nvt", "npt"):
            raise DeckError(f"fix {style} is not supported (nve, nvt, npt)")
        kw = dict(p_flag=[0, 0, 0], p_start=[0.0] * 3, p_stop=[0.0] * 3, p_damp=[1.0] * 3)
        i = 0
        while i < len(args):
            if args[i] == "temp":
                kw.update(t_start=float(args[i + 1]), t_stop=float(args[i + 2]), t_damp=float(args[i + 3]))
                i += 4
            elif args[i] in ("x", "y", "z") and style == "npt":
                d = "xyz".index(args[i])
                kw["p_flag"][d], kw["p_start"][d], kw["p_stop"][d], kw["p_damp"][d] = 1, float(args[i + 1]), float(args[i + 2]), float(args[i + 3])
                i += 4
            else:
                raise DeckError(f"fix {style}: keyword {args[i]} is not supported (temp, x, y, z)")
        if "t_start" not in kw:
            raise DeckError(f"fix {style} needs the temp keyword")
        self.fix = (a[0], style, kw)

    def cmd_unfix(self, a): self.fix = None

    # ------------------------------------------------------------------ state
    def _mass(self):
        m = self.masses.get(1, self.masses.get(0))
        if m is None:
            raise DeckError("Not all per-type masses are set")
        return m

    def _md(self):
        """(Re)create the device-resident driver for the current atoms, box and boundary."""
        if self.md is not None:
            return self.md
        import torch
        from .md import DomainMD, decompose
        if self.pair is None or self.x is None:
            raise DeckError("pair_coeff and read_data must come before this command")
        if self.world > 1:
            import torch.distributed as dist
            local = int(os.environ.get("LOCAL_RANK", 0))
            torch.cuda.set_device(local)
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        grid = decompose(self.world)
        req = getattr(self, "grid_request", None)
        if req and all(req) and req[0] * req[1] * req[2] == self.world:
            grid = tuple(req)
        boxlen = self.box[:, 1] - self.box[:, 0]
        xs = self.x - self.box[:, 0]
        cell = np.clip(np.floor(xs / (boxlen / np.array(grid))).astype(np.int64), 0, np.array(grid) - 1)
        owner = cell[:, 0] + cell[:, 1] * grid[0] + cell[:, 2] * grid[0] * grid[1]
        gid = np.nonzero(owner == self.rank)[0]
        periodic = tuple(b == "p" for b in self.boundary)
        wrap = tuple(b in ("m", "s") for b in self.boundary)
        self.md = DomainMD(self.pair, xs[gid], boxlen, grid=grid, rank=self.rank, mass=self._mass(), dt=self.dt, skin=self.skin,
                           periodic=periodic, shrink_wrap=wrap, type_local=self.types[gid], gid_local=gid)
        self.md.reneighbor()
        return self.md

    # ------------------------------------------------------------------ thermo
    def _thermo_line(self, vals):
        return " ".join(f"{int(vals[k]):9d}" if k == "step" else f"{vals[k]:14.8g}" for k in self.thermo_cols)

    def _thermo_header(self):
        return " ".join(f"{THERMO_HEAD[k]:>9s}" if k == "step" else f"{THERMO_HEAD[k]:>14s}" for k in self.thermo_cols)

    def _values(self, step, pe, ke_tensor, virial, box):
        nk = 1.6021765e6
        vol = box[0] * box[1] * box[2]
        ke = 0.5 * sum(ke_tensor[:3])
        ntot = len(self.x)
        p = [(ke_tensor[d] + virial[d]) / vol * nk for d in range(3)]
        return {"step": step, "temp": 2.0 * ke / ((3.0 * ntot - 3.0) * 8.617343e-5), "pe": pe, "ke": ke, "etotal": pe + ke,
                "lx": box[0], "ly": box[1], "lz": box[2], "press": sum(p) / 3.0, "vol": vol, "pxx": p[0], "pyy": p[1], "pzz": p[2]}

    # ------------------------------------------------------------------ minimize / run
    def cmd_minimize(self, a):
        md = self._md()
        st = md.minimize(float(a[0]), float(a[1]), int(a[2]), int(a[3]))
        self.last_minimize = st
        self.say("Minimization stats:")
        self.say(f"  Stopping criterion = {st['stopping_criterion']}")
        self.say("  Energy initial, next-to-last, final = ")
        self.say(f"    {st['energy_initial']:18.15g} {st['energy_next_to_last']:18.15g} {st['energy_final']:18.15g}")
        self.say(f"  Force two-norm initial, final = {st['fnorm_initial']:.8g} {st['fnorm_final']:.8g}")
        self.say(f"  Force max component initial, final = {st['fmax_initial']:.8g} {st['fmax_final']:.8g}")
        self.say(f"  Final line search alpha, max atom move = {st['alpha_final']:.8g} {st['max_atom_move']:.8g}")
        self.say(f"  Iterations, force evaluations = {st['iterations']} {st['evaluations']}")
        # atoms and box carry over to whatever follows (velocity, fix, run)
        import torch
        xs = torch.zeros((len(self.x), 3), dtype=torch.float64, device=md.dev)
        xs[md.gid] = md.x[: md.nlocal]
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(xs)
        self.x = xs.cpu().numpy() + self.box[:, 0]
        self.md = None

    def cmd_run(self, a):
        import time
        import torch
        nsteps = int(a[0])
        if self.fix is None:
            raise DeckError("run without a time-integration fix")
        md = self._md()
        _, style, kw = self.fix
        self.rows = []
        self.say(self._thermo_header())
        t0 = time.perf_counter()
        every = max(1, self.every if self.check else 1)
        base = md.nsteps
        if style == "nve":
            md.compute(eflag=True)
            log = md.run(nsteps, check_every=every, thermo_every=self.thermo_every)
            nan = float("nan")
            for s, pe, ke in log:                      # fix nve path evaluates no virial: pressure columns print nan
                b = list(md.wrap_hi - md.wrap_lo) if any(md.shrink_wrap) else list(md.box)
                v = {"step": self.step + (s - base), "temp": 2.0 * ke / ((3.0 * len(self.x) - 3.0) * 8.617343e-5), "pe": pe, "ke": ke,
                     "etotal": pe + ke, "lx": b[0], "ly": b[1], "lz": b[2], "press": nan, "vol": b[0] * b[1] * b[2], "pxx": nan, "pyy": nan, "pzz": nan}
                self.rows.append(v)
                self.say(self._thermo_line(v))
        else:
            md.fix_nh(kw["t_start"], kw["t_stop"], kw["t_damp"], p_flag=kw["p_flag"], p_start=kw["p_start"], p_stop=kw["p_stop"],
                      p_damp=kw["p_damp"], nsteps_ramp=nsteps)
            st = md.nh_state()
            pe0 = md.engvir[:1].clone()
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(pe0)
            v = self._values(self.step, float(pe0), list(st.ke_tensor[:]), list(st.virial[:]), [st.boxhi[d] - st.boxlo[d] for d in range(3)])
            if self.thermo_every:
                self.rows.append(v)
                self.say(self._thermo_line(v))
            for s, pe, ke, ext, T, p, b in md.run_nh(nsteps, check_every=every, thermo_every=self.thermo_every):
                vol = b[0] * b[1] * b[2]
                v = {"step": self.step + (s - base), "temp": T, "pe": pe, "ke": ke, "etotal": pe + ke,
                     "lx": b[0], "ly": b[1], "lz": b[2], "press": sum(p) / 3.0, "vol": vol, "pxx": p[0], "pyy": p[1], "pzz": p[2]}
                self.rows.append(v)
                self.say(self._thermo_line(v))
        torch.cuda.synchronize(md.dev)
        secs = time.perf_counter() - t0
        self.step += nsteps
        n = len(self.x)
        self.say(f"Loop time of {secs:.6g} on {self.world} procs for {nsteps} steps with {n} atoms")
        self.say(f"Performance: {86400.0 * nsteps * self.dt * 1e-3 / secs:.3f} ns/day, {nsteps / secs:.3f} timesteps/s, {n * nsteps / secs:.4g} atom-step/s")
        self.say(f"Neighbor list builds = {md.rebuilds}")
        self.loop_seconds = secs


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print(__doc__)
        return 2
    Deck().run_file(argv[0])
    return 0


if __name__ == "__main__":
    sys.exit(main())
