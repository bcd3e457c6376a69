"""Command line driver: same options and output as the reference's ``pyjess`` CLI
(src/pyjess/cli.py:85-283), batched.

The reference loads one query file at a time and runs `Jess.query` on it
(cli.py:257-278, optionally over a thread pool).  Here the query files are loaded with the
native parser and handed to the engine in batches through `Jess.query_many` (ONE launch
sequence per batch); the text written to stdout is byte-identical: one
`Hit.dump` block + blank line per hit, structures in list order, hits in the reference's
own iteration order.  `-i/--ignore-chain` selects ``ignore_chain="atoms"`` (the reference
passes the invalid value ``"all"`` there, cli.py:248, and fails).
"""
from __future__ import annotations

import argparse
import io
import sys
import typing
import warnings
from pathlib import Path

from . import __version__, sites
from ._jess import Jess
from ._model import Molecule, Template


def argument_parser(prog: str = "pyjess_b200") -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(
        prog=prog, add_help=False, formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        description="pyjess_b200 - PyJess's template queries on B200 GPUs (same options as the `pyjess` CLI).")
    parser.add_argument("-h", "--help", action="help", help="Show this help message and exit.")
    parser.add_argument("-V", "--version", action="version", version=f"pyjess_b200 {__version__}",
                        help="Show the version number and exit.")
    parser.add_argument("-j", "--jobs", type=int, default=1,
                        help="Threads reading the query files of a batch (the search itself runs on the GPU).")
    parser.add_argument("--batch", type=int, default=256, help="Query structures per engine call.")
    group = parser.add_argument_group("Mandatory Parameters")
    group.add_argument("-T", "--templates", type=Path, required=True, help="The path to the template list file.")
    group.add_argument("-Q", "--queries", type=Path, required=True, help="The path to the query list file.")
    group.add_argument("-R", "--rmsd", type=float, required=True, help="The RMSD threshold.")
    group.add_argument("-D", "--distance-cutoff", type=float, required=True, help="The distance-cutoff.")
    group.add_argument("-M", "--maximum-distance", type=float, required=True,
                       help="The maximum allowed template/query atom distance after adding the global distance "
                            "cutoff and the individual atom distance cutoff of the template file.")
    group = parser.add_argument_group("Flags")
    group.add_argument("-n", "--no-transform", action="store_false", dest="transform",
                       help="Do not transform coordinates of hit into the template coordinate frame")
    group.add_argument("-f", "--filenames", action="store_true", help="Show PDB filenames in progress on stderr")
    group.add_argument("-i", "--ignore-chain", action="store_true",
                       help="Include matches composed of residues belonging to multiple chains (if template is "
                            "single-chain) or matches with residues from a single chain (if template has residues "
                            "from multiple chains).")
    group.add_argument("--ignore-res-chain", action="store_true",
                       help="Include matches composed of residues belonging to multiple chains but still enforce "
                            "all atoms of a residue to be part of the same chain.")
    group.add_argument("-q", "--query-filename", action="store_true",
                       help="Write filename of query instead of PDB ID from HEADER")
    group.add_argument("-e", "--ignore-endmdl", action="store_true",
                       help="Parse atoms from all models separated by ENDMDL (use with care).")
    group.add_argument("-c", "--max-candidates", type=int, default=None,
                       help="Set a maximum number of candidates to return by template.")
    group.add_argument("--no-reorder", action="store_false", dest="reorder",
                       help="Disable template atom reordering in the matching process: results come exactly in the "
                            "order of the original Jess.")
    group.add_argument("-b", "--best-match", action="store_true",
                       help="Return only the best match for each template/query pair.")
    group.add_argument("--best-sites", choices=("all", "best"), default=None,
                       help="Post-filter of the reference's vendor/jess/filter_output.py: keep the lowest log(E) hit "
                            "per distinct list of matched residues of every template/query pair and drop self "
                            "matches ('all'), or only the best of those per pair ('best'). Applied per batch.")
    return parser


def main(argv: typing.Optional[typing.List[str]] = None, stdout: typing.TextIO = sys.stdout,
         stderr: typing.TextIO = sys.stderr) -> int:
    args = argument_parser().parse_args(argv)
    ignore_chain = "atoms" if args.ignore_chain else "residues" if args.ignore_res_chain else None
    try:
        with args.templates.open() as f:
            jess = Jess([Template.load(n, id=n) for n in filter(None, map(str.strip, f))])
        with args.queries.open() as f:
            filenames = [n for n in map(str.strip, f) if n]
        for start in range(0, len(filenames), max(1, args.batch)):
            mols = []
            names = filenames[start:start + max(1, args.batch)]
            loaded = Molecule.load_many(names, jobs=max(1, args.jobs), ids=names if args.query_filename else None,
                                        ignore_endmdl=args.ignore_endmdl)
            for filename, mol in zip(names, loaded):
                if mol.id is None:
                    warnings.warn(f"No id in molecule {filename}. Using filename instead.")
                    mol = Molecule(mol, id=filename, date=mol.date, name=mol.name)
                if args.filenames:
                    print(filename, file=stderr)
                mols.append(mol)
            queries = jess.query_many(mols, args.rmsd, args.distance_cutoff, args.maximum_distance,
                                      max_candidates=args.max_candidates, ignore_chain=ignore_chain,
                                      reorder=args.reorder, best_match=args.best_match)
            if args.best_sites is not None and args.max_candidates is None:
                sites.write_hits(sites.best_sites(queries, all_instances=args.best_sites == "all"), stdout,
                                 transform=args.transform)
                continue
            text = io.StringIO() if args.best_sites is not None else stdout
            for query in queries:
                for hit in query:
                    hit.dump(text, format="pdb", transform=args.transform)
                    text.write("\n")
            if args.best_sites is not None:    # max_candidates: the replayed stream is filtered as text
                text.seek(0)
                sites.filter_text(text, stdout, all_instances=args.best_sites == "all")
    except Exception as err:  # same contract as the reference CLI: message on stderr, errno or 1
        print("Error: {}".format(err), file=stderr)
        return getattr(err, "errno", None) or 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
