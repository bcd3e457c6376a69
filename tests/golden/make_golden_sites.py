"""Golden output of the reference's post-filter (vendor/jess/filter_output.py).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_sites.py
Feeds the committed command-line goldens (cli_*.txt, written by the reference's `pyjess`
CLI, see make_golden.py) through the UNMODIFIED script: `main(True)` is what the script does
when executed, `main(False)` its other mode (one hit per (target, template) run).
"""
import contextlib
import importlib.util
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = "/root/reference/vendor/jess/filter_output.py"

spec = importlib.util.spec_from_file_location("filter_output", SCRIPT)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

for name in ("cli_default", "cli_flags", "cli_maxcand"):
    with open(os.path.join(HERE, name + ".txt")) as f:
        text = f.read()
    for tag, flag in (("all", True), ("best", False)):
        sys.stdin = io.StringIO(text)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            mod.main(flag)
        sys.stdin = sys.__stdin__
        path = os.path.join(HERE, f"sites_{name}_{tag}.txt")
        with open(path, "w") as f:
            f.write(buf.getvalue())
        print(path, text.count("REMARK"), "->", buf.getvalue().count("REMARK"))

# adversarial synthetic stream: score ties, self matches, pairs that come back after another
# pair, HETATM records, residues that differ only in chain, permuted and repeated residues
import random

rng = random.Random(20261018)
lines = []
residues = [("SER", c, n) for c in ("A", "B", " ") for n in (7, 8, 9, 10)] + [("HIS", "A", 7), ("ASP", "AB", 1234)]
pairs = [("1ABC", "tplA"), ("1ABC", "tplB"), ("tplA", "tplA"), ("2XYZ", "tplA"), ("1ABC", "tplA"), ("2XYZ", "2XYZ"),
         ("3QRS", "tplC")]
serial = 0
for target, template in pairs:
    for _ in range(rng.randint(30, 70)):
        rmsd = rng.choice([0.5, 0.75, 1.0, 1.004, 1.25])
        score = rng.choice([-1.5, -1.25, 0.0, 0.5, 0.51, 2.0])
        lines.append("REMARK %s %5.3f %s Det= 1.0 log(E)~ %4.2f" % (target, rmsd, template, score))
        for _ in range(rng.randint(1, 3)):
            res = rng.choice(residues[:6] if rng.random() < 0.8 else residues)
            for _ in range(rng.randint(1, 2)):
                serial += 1
                rec = rng.choice(["ATOM  ", "ATOM  ", "HETATM"])
                lines.append("%s%5i  CA  %-3s%2s%4i    %8.3f%8.3f%8.3f  1.00 20.00           C" % (
                    rec, serial % 100000, res[0], res[1], res[2], rng.uniform(-9, 9), rng.uniform(-9, 9), rng.uniform(-9, 9)))
        lines.append("ENDMDL")
        lines.append("")
text = "\n".join(lines) + "\n"
with open(os.path.join(HERE, "sites_synthetic_in.txt"), "w") as f:
    f.write(text)
for tag, flag in (("all", True), ("best", False)):
    sys.stdin = io.StringIO(text)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mod.main(flag)
    sys.stdin = sys.__stdin__
    with open(os.path.join(HERE, f"sites_synthetic_{tag}.txt"), "w") as f:
        f.write(buf.getvalue())
    print("synthetic", tag, text.count("REMARK"), "->", buf.getvalue().count("REMARK"))
