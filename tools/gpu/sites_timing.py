"""Cost of the best-hit-per-site filter on the cfg3 hit table: `sites.best_sites`
over the engine's table against the reference's route (dump every hit as text, parse and filter
it: vendor/jess/filter_output.py, here its byte-identical restatement `sites.filter_text`)."""
import io
import json
import sys
import time

from pyjess_b200 import Jess, sites, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
tpl, mols, thr, _ = synth.workload("cfg3", scale=scale)
jess = Jess(tpl)
list(jess.query_many(mols[:8], *thr)[0])                     # warm-up (context, tables)
t0 = time.perf_counter()
queries = jess.query_many(mols, *thr)
n_hits = sum(len(q.hit_table()[1]) for q in queries)
t1 = time.perf_counter()
kept = sites.best_sites(queries, all_instances=True)
t2 = time.perf_counter()
sample = max(1, len(mols) // 10)
text = io.StringIO()
n_sample = 0
for q in jess.query_many(mols[:sample], *thr):
    for hit in q:
        hit.dump(text)
        text.write("\n")
        n_sample += 1
t3 = time.perf_counter()
text.seek(0)
out = io.StringIO()
n_text = sites.filter_text(text, out, all_instances=True)
t4 = time.perf_counter()
first = sites.best_sites(jess.query_many(mols[:sample], *thr), all_instances=True)
print(json.dumps({
    "workload": f"cfg3 x{scale}", "pairs": len(tpl) * len(mols), "hits": n_hits, "kept": len(kept),
    "query_many_s": round(t1 - t0, 4), "best_sites_s": round(t2 - t1, 4),
    "us_per_hit_table_route": round(1e6 * (t2 - t1) / max(1, n_hits), 3),
    "text_route_sample": {"structures": sample, "hits": n_sample, "dump_s": round(t3 - t2, 4),
                          "filter_text_s": round(t4 - t3, 4), "kept": n_text,
                          "us_per_hit": round(1e6 * (t4 - t2) / max(1, n_sample), 3),
                          "same_selection_as_table_route": n_text == len(first)},
}))
