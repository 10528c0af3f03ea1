"""Differential fuzz of the WHOLE pipeline on the emulated device: random Filter/Project/Agg plans over random typed,
nullable inputs, random batch sizes and confs — library (emulated device) vs the numpy oracle.
usage: python tools/emu/fuzz_pipeline.py [seed] [cases]"""
import decimal
import os
import subprocess
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pyarrow as pa


def main(seed, cases):
    out = os.environ.get("B200Q_EMU_DIR") or os.path.join(os.environ.get("TMPDIR", "/tmp"), "b200q_emu")
    lib = subprocess.run([os.path.join(HERE, "build_lib.sh"), out], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    from blaze_b200 import native
    native.LIB_PATH = lib; native.lib = native._load()
    from blaze_b200 import exprs as E, plans as PL, types as T
    from oracle import blaze_oracle as O
    import helpers as H

    INT_TYPES = [(pa.int64(), T.int64, np.int64), (pa.int32(), T.int32, np.int32), (pa.int16(), T.int16, np.int16), (pa.int8(), T.int8, np.int8)]
    failures = 0
    for case in range(cases):
        rng = np.random.default_rng(seed * 100003 + case)
        n = int(rng.choice([0, 1, 7, 33, 257, 1000, 2500, 4099]))
        def nulls(vals, pa_t, frac): return pa.array(vals, type=pa_t, mask=(rng.random(len(vals)) < frac) if frac > 0 and len(vals) else None)
        nf = float(rng.choice([0.0, 0.0, 0.05, 0.4]))
        kt = INT_TYPES[rng.integers(0, 4)]; k2t = INT_TYPES[rng.integers(0, 4)]
        krange = int(rng.choice([3, 40, 100, 2000])); k2range = int(rng.choice([2, 5, 100]))
        lo = int(rng.choice([-50, 0, 10])) if kt[2] == np.int8 else int(rng.choice([-50, 0, 10**4]))
        k = (lo + rng.integers(0, min(krange, 100 if kt[2] == np.int8 else krange), n)).astype(kt[2])
        k2 = rng.integers(0, min(k2range, 100), n).astype(k2t[2])
        v = rng.integers(-2**50, 2**50, n, dtype=np.int64)
        w = rng.integers(-30000, 30000, n).astype(np.int32)
        x = rng.normal(0, 1e3, n)
        dec = [decimal.Decimal(int(r)).scaleb(-2) for r in rng.integers(-10**12, 10**12, n)]
        f = rng.integers(0, 20, n).astype(np.int64)
        rb = pa.RecordBatch.from_arrays([nulls(k, kt[0], nf / 4), nulls(k2, k2t[0], nf / 4), nulls(v, pa.int64(), nf), nulls(w, pa.int32(), nf), nulls(x, pa.float64(), nf),
                                         pa.array(dec, type=pa.decimal128(17, 2), mask=(rng.random(n) < nf) if nf > 0 and n else None), nulls(f, pa.int64(), nf / 2)],
                                        names=["k", "k2", "v", "w", "x", "d", "f"])
        bs = int(rng.choice([1, 100, 1000, 10000]))
        bs = max(bs, (n + 63) // 64)                                   # at most 64 pushes per case: every emulated launch spawns OS threads
        batches = H.split_batches(rb, max(1, min(bs, max(n, 1)))) if n else [rb]
        leaf = PL.MemoryExec.from_arrow(batches, rb.schema); ins = leaf.schema()
        preds = []
        if rng.random() < 0.6: preds.append(E.BinaryExpr(E.Column("f"), str(rng.choice(["Lt", "GtEq", "NotEq"])), E.Literal(int(rng.integers(0, 20)), T.int64)))
        if rng.random() < 0.3: preds.append(E.BinaryExpr(E.Literal(int(rng.integers(-20000, 20000)), T.int32), "LtEq", E.Column("w")))
        if rng.random() < 0.15: preds.append(E.IsNotNull(E.Column("x")))
        conf = native.default_conf(staging_rows=int(rng.choice([0, 1 << 20])), max_launch_rows=int(rng.choice([1 << 16, 1 << 27])),
                                   agg_dense_keys=int(rng.random() < 0.7), force_generic_kernels=int(rng.random() < 0.15), agg_hot_key_cache=int(rng.random() < 0.3))
        kind = rng.choice(["filter", "project", "agg", "agg", "agg2"])
        desc = f"case {case}: n={n} nf={nf} kt={kt[0]} k2t={k2t[0]} bs={bs} kind={kind} preds={len(preds)} dense={conf.agg_dense_keys} generic={conf.force_generic_kernels} hot={conf.agg_hot_key_cache} staging={conf.staging_rows}"
        try:
            ob = H.oracle_batches(batches)
            if kind == "filter":
                if not preds: preds = [E.BinaryExpr(E.Column("v"), "Gt", E.Literal(0, T.int64))]
                plan = PL.FilterExec(preds, leaf)
                got = PL.collect(plan, conf); exp = O.FilterExec(preds, ins).execute(ob)
                H.assert_same_rows_ordered(got, exp, plan.schema())
            elif kind == "project":
                projs = [(E.Column("v"), "v"), (E.BinaryExpr(E.Column("v"), "Plus", E.Cast(E.Column("w"), T.int64)), "s"), (E.BinaryExpr(E.Column("x"), "Multiply", E.Literal(2.0, T.float64)), "x2"),
                         (E.Case(None, [(E.BinaryExpr(E.Column("f"), "Lt", E.Literal(5, T.int64)), E.Column("v"))], E.Literal(None, T.int64)), "c")]
                plan = PL.ProjectExec(projs, PL.FilterExec(preds, leaf) if preds else leaf)
                got = PL.collect(plan, conf); exp = O.ProjectExec(projs, ins, preds).execute(ob)
                H.assert_same_rows_ordered(got, exp, plan.schema())
            else:
                nkeys = int(rng.choice([0, 1, 1, 2]))
                groupings = [E.GroupingExpr("k", E.Column("k"))][:nkeys] + ([E.GroupingExpr("k2", E.Column("k2"))] if nkeys == 2 else [])
                pool = [("sv", E.AGG_SUM, "v", T.int64), ("cv", E.AGG_COUNT, "v", T.int64), ("sw", E.AGG_SUM, "w", T.int64), ("c1", E.AGG_COUNT, None, T.int64),
                        ("mnv", E.AGG_MIN, "v", T.int64), ("mxw", E.AGG_MAX, "w", T.int32), ("av", E.AGG_AVG, "v", T.float64), ("sx", E.AGG_SUM, "x", T.float64),
                        ("sd", E.AGG_SUM, "d", T.decimal128(27, 2)), ("ad", E.AGG_AVG, "d", T.decimal128(21, 6)), ("mxx", E.AGG_MAX, "x", T.float64)]
                simple = rng.random() < 0.6                         # the specialised kernels take 1-2 add-class aggregates
                cand = pool[:4] if simple else pool
                pick = [cand[i] for i in sorted(rng.choice(len(cand), size=int(rng.integers(1, 3 if simple else 5)), replace=False))]
                def mk(mode, schema, final):
                    out = []
                    for nm, fn, col, rt in pick:
                        ch = [E.Literal(1, T.int64)] if col is None else ([E.placeholder(E.Column(col).data_type(ins))] if final else [E.Column(col)])
                        out.append(E.AggExpr(nm, mode, PL.create_agg(fn, ch, schema, rt)))
                    return out
                child = PL.FilterExec(preds, leaf) if preds else leaf
                partial = PL.AggExec(PL.HashAgg, groupings, mk(E.PARTIAL, ins, False), bool(rng.random() < 0.5), child)
                o_child = O.FilterExec(preds, ins).execute(ob) if preds else ob
                o_partial = O.AggExec(E.HASH_AGG, groupings, mk(E.PARTIAL, ins, False), False, ins)
                fcols = tuple(i + nkeys for i, p in enumerate(pick) if p[3] == T.float64)
                if kind == "agg":
                    got = PL.collect(partial, conf); exp = o_partial.execute(o_child)
                    # the frozen Binary column is compared byte for byte unless it embeds fp64 sums (order of the atomic adds)
                    if fcols:
                        continue
                    H.assert_multiset_equal(got, exp)
                else:
                    final = PL.AggExec(PL.HashAgg, groupings, mk(E.FINAL, partial.schema(), True), False, partial)
                    o_final = O.AggExec(E.HASH_AGG, groupings, mk(E.FINAL, o_partial.schema, True), False, o_partial.schema)
                    got = PL.collect(final, conf); exp = o_final.execute(o_partial.execute(o_child))
                    H.assert_multiset_equal(got, exp, fcols)
        except Exception:
            failures += 1
            print("FAIL", desc); traceback.print_exc(limit=3); sys.stdout.flush()
            if failures >= 5: break
            continue
        if case % 10 == 0: print("ok  ", desc); sys.stdout.flush()
    print(f"{cases} cases, {failures} failures")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 50))
