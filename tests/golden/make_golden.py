"""Generate the committed golden fixtures: `python tests/golden/make_golden.py` (CPU only, deterministic).
Expected outputs come from the numpy oracle (oracle/blaze_oracle.py), itself pinned to the reference's KATs."""
import json, os, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import numpy as np
import pyarrow as pa

from oracle import blaze_oracle as O
import cases


def oracle_run(plan, batches):
    """execute a blaze_b200.plans tree with the oracle's operators"""
    from blaze_b200 import plans as PL
    if isinstance(plan, PL.MemoryExec):
        return [O.batch_from_arrow(b) for b in batches], plan.schema()
    child, ins = oracle_run(plan.input, batches)
    if isinstance(plan, PL.FilterExec):
        return O.FilterExec(plan.predicates, ins).execute(child), ins
    if isinstance(plan, PL.ProjectExec):
        op = O.ProjectExec(plan.exprs, ins, [])
        return op.execute(child), op.schema
    op = O.AggExec(plan.exec_mode, plan.groupings, plan.aggs, plan.supports_partial_skipping, ins)
    return op.execute(child), op.schema


def expected_table(case):
    from blaze_b200 import plans as PL
    batches = [case.rb.slice(i, min(case.batch_rows, case.rb.num_rows - i)) for i in range(0, case.rb.num_rows, case.batch_rows)]
    leaf = PL.MemoryExec.from_arrow(batches, case.rb.schema)
    out, schema = oracle_run(case.build(leaf), batches)
    return O.batch_to_arrow(O.concat_batches(schema, out))


def write_ipc(path, rb):
    with pa.OSFile(path, "wb") as f, pa.ipc.new_file(f, rb.schema) as w:
        w.write_batch(rb)


def main():
    manifest = {}
    for c in cases.all_cases():
        exp = expected_table(c)
        write_ipc(os.path.join(HERE, c.name + ".in.arrow"), c.rb)
        write_ipc(os.path.join(HERE, c.name + ".out.arrow"), exp)
        manifest[c.name] = {"rows_in": c.rb.num_rows, "rows_out": exp.num_rows, "ordered": c.ordered}
    a, b, valid_b, nparts = cases.murmur3_case()
    cols = [O.Col(O.T.int64, a, np.ones(len(a), bool)), O.Col(O.T.int32, b, valid_b)]
    h = O.create_murmur3_hashes(cols, len(a), 42)
    np.savez_compressed(os.path.join(HERE, "murmur3_partition.npz"), a=a, b=b, valid_b=valid_b, hashes=h,
                        **{"p%d" % n: O.partition_ids(h, n) for n in nparts})
    json.dump(manifest, open(os.path.join(HERE, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
