"""Early-stopping rules with the reference's semantics (elliot/recommender/early_stopping.py:8-134):
monitor `loss` (mode min) or a validation metric[@k] (mode max); stop when, over the last
`patience`+1 observations, every consecutive pair got worse (or improved by no more than
min_delta / rel_delta, or is still on the wrong side of `baseline`).  Host control flow only."""
from types import SimpleNamespace


class EarlyStopping:
    def __init__(self, ns: SimpleNamespace, validation_metric, validation_k, cutoffs, simple_metrics):
        opts = dict(ns.__dict__)
        self.active = bool(opts)
        self.validation_metric, self.validation_k = validation_metric, validation_k
        if not self.active:
            return
        self.patience = opts.get("patience", 0)
        self.monitor = opts.get("monitor", validation_metric)
        self.verbose = opts.get("verbose", False)
        self.metric = None
        if self.monitor == "loss":
            self.mode = "min"
        else:
            self.mode = "max"
            name, _, k = self.monitor.partition("@")
            if name.lower() not in [m.lower() for m in simple_metrics]:
                raise Exception("Early stopping metric must be in the list of simple metrics")
            self.metric_k = int(k) if k else validation_k
            if self.metric_k not in cutoffs:
                raise Exception("Validation cutoff must be in general cutoff values")
            self.metric = name
        if opts.get("mode", "auto") in ("min", "max"):
            self.mode = opts["mode"]
        self.rules = {k: opts[k] for k in ("min_delta", "rel_delta", "baseline") if k in opts}

    def _worse(self, a, b):
        """a = newer-side observation, b = the one after it in the scan order of the reference."""
        if b > a:
            return True
        r = self.rules
        if "min_delta" in r and (a - b) <= r["min_delta"]:
            return True
        if "rel_delta" in r and (a - b) <= a * r["rel_delta"]:
            return True
        if "baseline" in r:
            return a >= r["baseline"] if self.mode == "min" else a <= r["baseline"]
        return False

    def stop(self, losses, results):
        if not self.active:
            return False
        obs = list(losses) if self.metric is None else \
            [r[self.metric_k]["val_results"][self.metric] for r in results]
        if len(obs) <= self.patience:
            return False
        window = obs[:-(2 + self.patience):-1]       # newest first, patience+1 values
        if self.mode == "min":
            window = window[::-1]
        checks = [self._worse(window[p], window[p + 1]) for p in range(len(window) - 1)]
        return bool(checks) and all(checks)

    def __str__(self):
        return ", ".join(f"{k}: {v}" for k, v in self.__dict__.items())
