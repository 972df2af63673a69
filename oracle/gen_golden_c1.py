#!/usr/bin/env python
"""Mint the C1-scale golden (BASELINE.json configs[0]: BPRMF d=64 on an ML-1M-shaped matrix, reference CPU path via
run_experiment, config_files/sample_hello_world.yml:2-10 shape) by running the UNMODIFIED reference end to end:

    elliot.run.run_experiment(<yaml>)   with a `BPRMF:` block (factors 64, 10 epochs, the reference's default
                                        hyper-parameters, seed 42), `strategy: dataset`, `random_subsampling 0.2`

on the synthetic 6 040 x 3 706 / ~1.0 M-rating matrix of elliot_b200/synth_c1.py (MovieLens-1M itself is not shipped and
there is no network).  tensorflow/hyperopt are stubbed for import only (oracle/ref_stubs.py); BPRMF is pure NumPy.
Captured (tests/golden/bprmf_c1.npz): the reference Evaluator's test metrics after EVERY epoch (a pass-through wrapper
around Evaluator.eval records them — the reference only logs them), the recommendation lists the reference stored
(`save_recs`), the dataset checksum.  Build container only (~6 min of CPU); the GPU tests read the .npz.

    python oracle/gen_golden_c1.py [--epochs 10]
"""
import argparse
import glob
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_stubs  # noqa: E402
from elliot_b200 import synth_c1  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "bprmf_c1.npz")
METRICS = ["nDCG", "HR", "Precision", "Recall"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--factors", type=int, default=64)
    args = ap.parse_args()
    ref_stubs.install()
    tmp = tempfile.mkdtemp(prefix="c1_golden_")
    tsv = os.path.join(tmp, "dataset.tsv")
    checksum = synth_c1.write_tsv(tsv)
    logcfg = ref_stubs.write_logger_config(os.path.join(tmp, "logger_config.yml"))
    cfg = os.path.join(tmp, "cfg.yml")
    with open(cfg, "w") as fh:
        fh.write(synth_c1.yaml_text(tsv, tmp, "BPRMF", args.epochs, args.factors, extra=f"  path_logger_config: {logcfg}\n"))

    from elliot.evaluation.evaluator import Evaluator
    per_epoch = []
    orig_eval = Evaluator.eval

    def recording_eval(self, recommendations):                      # pass-through: records what the reference computed
        res = orig_eval(self, recommendations)
        k = list(res.keys())[0]
        per_epoch.append([float(res[k]["test_results"][m]) for m in METRICS])
        print(f"epoch {len(per_epoch)}: " + " ".join(f"{m}={v:.6f}" for m, v in zip(METRICS, per_epoch[-1])), flush=True)
        return res
    Evaluator.eval = recording_eval
    from elliot.run import run_experiment
    t0 = time.time()
    run_experiment(cfg)
    dt = time.time() - t0
    Evaluator.eval = orig_eval
    rec_files = sorted(glob.glob(os.path.join(tmp, "recs", "*.tsv")))
    assert rec_files, "the reference stored no recommendation file"
    rec = np.loadtxt(rec_files[-1], delimiter="\t")
    users, first = np.unique(rec[:, 0].astype(np.int64), return_index=True)
    keep = 400                                                      # lists of the first 400 users (by public id)
    sel = np.isin(rec[:, 0].astype(np.int64), users[:keep])
    np.savez_compressed(OUT, metrics=np.array(METRICS), per_epoch=np.array(per_epoch), epochs=args.epochs, factors=args.factors,
                        rec_users=rec[sel, 0].astype(np.int64), rec_items=rec[sel, 1].astype(np.int64), rec_scores=rec[sel, 2],
                        rec_file=os.path.basename(rec_files[-1]), checksum=np.uint64(checksum), n_rec_users=len(users),
                        reference_seconds=dt)
    print(f"wrote {OUT}: {len(per_epoch)} epochs, {int(sel.sum())} rec rows, reference run {dt:.0f} s")
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
