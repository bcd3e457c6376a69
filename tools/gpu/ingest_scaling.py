"""Structure ingest on the host cores of the box: `Molecule.load_many` over the GIL-free native
parsers (jb_pdb_parse / jb_cif_parse) against one `Molecule.load` per file."""
import json
import os
import sys
import time
import warnings

sys.path.insert(0, ".")
from pyjess_b200 import Molecule  # noqa: E402

warnings.simplefilter("ignore")
data = os.path.join("tests", "data")
out = {"cores": len(os.sched_getaffinity(0))}
for label, names in (("pdb", ["1AMY.pdb", "pdb1lnb.ent", "pdb1a0p.ent"]), ("cif", ["1AMY.cif"])):
    files = [os.path.join(data, n) for n in names] * (600 // len(names))
    atoms = sum(len(m) for m in Molecule.load_many(files[:len(names)], jobs=1)) * (len(files) // len(names))
    row = {"files": len(files), "atoms": atoms}
    for jobs in (1, 2, 4, 8, 16):
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            Molecule.load_many(files, jobs=jobs)
            best = min(best, time.perf_counter() - t)
        row[f"jobs{jobs}_ms_per_file"] = round(1e3 * best / len(files), 4)
    row["us_per_atom_1_thread"] = round(1e6 * row["jobs1_ms_per_file"] * 1e-3 * len(files) / atoms, 3)
    out[label] = row
print(json.dumps(out))
