"""Run the reference's own input deck (tests/golden/in.st_test, verbatim) with meng_zhang_b200.deck and write a LAMMPS-style
log.  The data file is re-created from tests/golden/fe_st.npz (the reference's fe_st.dat, numbers only), the potential
from tests/golden/fe_potential.json.   python scripts/run_reference_deck.py profiles/r1b_in.st_test.log"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from meng_zhang_b200.deck import Deck  # noqa: E402
from meng_zhang_b200.structures import write_lammps_data  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else None
rank = int(os.environ.get("RANK", 0))
td = os.path.join(tempfile.gettempdir(), "annp_b200_deck_run")
if rank == 0:
    os.makedirs(td, exist_ok=True)
    z = np.load(os.path.join(util.GOLDEN, "fe_st.npz"))
    write_lammps_data(os.path.join(td, "fe_st.dat"), z["x"], z["box"][:, 1], np.ones(len(z["x"]), dtype=np.int32), ntypes=1)
    util.write_fe_potential(os.path.join(td, "fe_annp_potential_2.ann"))
    shutil.copy(os.path.join(util.GOLDEN, "in.st_test"), os.path.join(td, "in.st_test"))
if int(os.environ.get("WORLD_SIZE", 1)) > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    dist.barrier()
fp = open(out_path, "w") if (out_path and rank == 0) else sys.stdout
Deck(out=fp).run_file(os.path.join(td, "in.st_test"))
