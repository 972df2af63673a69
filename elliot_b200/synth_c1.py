"""Synthetic data generator (no model arithmetic): the MovieLens-1M-SHAPED interaction file used for BASELINE.json configs[0]
(sample_hello_world.yml:2-10 names data/movielens_1m/dataset.tsv, which is not shipped and cannot be downloaded here).

6 040 users x 3 706 items, ~1.0 M ratings: per-user counts from a clipped log-normal (min 20 like ML-1M's filter), items
drawn without replacement from a Zipf-like popularity, ratings 1..5, rows shuffled, `user\\titem\\trating\\ttimestamp`
like the reference's loader expects (dataset.py:75-80).  Deterministic for a given numpy; the checksum of the rows is
stored in the golden so that a different numpy stream fails loudly instead of silently changing the data.
"""
import zlib

import numpy as np

N_USERS, N_ITEMS, SEED = 6040, 3706, 20240924


def rows(seed=SEED):
    g = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, N_ITEMS + 1) ** 0.9
    pop /= pop.sum()
    counts = np.clip(g.lognormal(np.log(110.0), 0.95, N_USERS), 20, 2300).astype(np.int64)
    us, its = [], []
    for u in range(N_USERS):
        c = int(counts[u])
        it = g.choice(N_ITEMS, size=c, replace=False, p=pop)
        us.append(np.full(c, u + 1, np.int64)); its.append(it.astype(np.int64) + 1)
    u = np.concatenate(us); i = np.concatenate(its)
    r = g.integers(1, 6, size=u.size).astype(np.int64)
    perm = g.permutation(u.size)
    return u[perm], i[perm], r[perm]


def checksum(u, i, r):
    return zlib.crc32(np.stack([u, i, r], 1).astype(np.int64).tobytes())


def write_tsv(path, seed=SEED):
    u, i, r = rows(seed)
    with open(path, "w") as fh:
        fh.write("".join(f"{a}\t{b}\t{float(c)}\t{t}\n" for t, (a, b, c) in enumerate(zip(u.tolist(), i.tolist(), r.tolist()))))
    return checksum(u, i, r)


def yaml_text(tsv, out_dir, model_key, epochs, factors, extra="", model_extra="", seed=42):
    """The reference's YAML layout (sample_hello_world.yml:1-19 with a `BPRMF:` block, BPRMF.py:43-56 keys)."""
    return f"""experiment:
  dataset: c1_synth
  data_config:
    strategy: dataset
    dataset_path: {tsv}
  splitting:
    test_splitting:
      strategy: random_subsampling
      test_ratio: 0.2
  top_k: 10
  evaluation:
    simple_metrics: [nDCG, HR, Precision, Recall]
  path_output_rec_result: {out_dir}/recs
  path_output_rec_weight: {out_dir}/weights
  path_output_rec_performance: {out_dir}/performance
  path_log_folder: {out_dir}/log
{extra}  models:
    {model_key}:
      meta:
        save_recs: True
        verbose: False
      epochs: {epochs}
      factors: {factors}
      lr: 0.05
      bias_regularization: 0
      user_regularization: 0.0025
      positive_item_regularization: 0.0025
      negative_item_regularization: 0.00025
      seed: {seed}
{model_extra}"""
