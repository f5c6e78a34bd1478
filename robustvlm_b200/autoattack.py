"""AutoAttack orchestrator over the CUDA APGD engine.

  AutoAttack          autoattack/autoattack.py:12-345   (same ctor / run_standard_evaluation /
                                                         clean_accuracy / set_version)
Only the APGD family is built (apgd-ce, apgd-dlr, apgd-t) -- the attacks the reference's CLIP
evaluation runs (`CLIP_eval/clip_robustbench.py:289-300`, version='custom',
attacks_to_run=['apgd-ce', 'apgd-t']).  FAB and Square are out of scope and
raise NotImplementedError when asked for; the 'standard' / 'plus' presets therefore keep
their APGD members and drop the rest, with a note on the logger.
The evaluation state file (`EvaluationState`, autoattack/state.py: resume an interrupted
evaluation) is kept in the reference's JSON layout; randomised-defence checks and TF adapters
of the reference are control plane and not rebuilt.
"""
from __future__ import annotations

import ast
import json
import math
import time
import warnings
from datetime import datetime
from pathlib import Path

import torch

from .attacks import APGDAttack, APGDAttack_targeted

_APGD_FAMILY = ("apgd-ce", "apgd-dlr", "apgd-t")
_NOT_BUILT = ("fab", "fab-t", "square")


class EvaluationState:
    """autoattack/state.py:11-89 -- which attacks have run and which points are still robust, saved
    as JSON so `run_standard_evaluation(state_path=...)` can resume.  Same keys as the reference's
    file.  (The reference writes `_attacks_to_run` through `json.dump(default=str)`, i.e. as the
    repr of a set, and then compares that string with a set on resume -- `autoattack.py:89-91`
    always raises; here the set is written as a sorted list and both forms are read.)"""

    SAVE_TIMEOUT = 60

    def __init__(self, attacks_to_run, path=None):
        self._attacks_to_run = set(attacks_to_run)
        self.path = Path(path) if path is not None else None
        self._run_attacks = set()
        self._robust_flags = None
        self._last_saved = datetime(1, 1, 1)
        self._clean_accuracy = float("nan")

    def to_disk(self, force=False):
        since = (datetime.now() - self._last_saved).total_seconds()
        if self.path is None or (since < self.SAVE_TIMEOUT and not force):
            return
        self._last_saved = datetime.now()
        d = {"_attacks_to_run": sorted(self._attacks_to_run), "path": str(self.path),
             "_run_attacks": sorted(self._run_attacks),
             "_robust_flags": (self._robust_flags.cpu().tolist()
                               if self._robust_flags is not None else None),
             "_last_saved": self._last_saved.isoformat(), "_SAVE_TIMEOUT": self.SAVE_TIMEOUT,
             "_clean_accuracy": self._clean_accuracy}
        tmp = self.path.with_suffix(self.path.suffix + ".tmp")
        with tmp.open("w") as f:
            json.dump(d, f)
        tmp.replace(self.path)

    @classmethod
    def from_disk(cls, path):
        path = Path(path)
        with path.open("r") as f:
            d = json.load(f)
        atr = d["_attacks_to_run"]
        if isinstance(atr, str):                 # a file written by the reference itself
            atr = ast.literal_eval(atr)
        st = cls(atr, path=d.get("path", path))
        if Path(d.get("path", path)) != path:
            warnings.warn(UserWarning("The given path is different from the one found in the state file."))
        st._run_attacks = set(d["_run_attacks"])
        if d.get("_robust_flags") is not None:
            st._robust_flags = torch.tensor(d["_robust_flags"], dtype=torch.bool)
        st._last_saved = datetime.fromisoformat(d["_last_saved"])
        st._clean_accuracy = float(d.get("_clean_accuracy", float("nan")))
        return st

    @property
    def robust_flags(self):
        return self._robust_flags

    @robust_flags.setter
    def robust_flags(self, flags):
        self._robust_flags = flags
        self.to_disk(force=True)

    @property
    def run_attacks(self):
        return self._run_attacks

    def add_run_attack(self, attack):
        self._run_attacks.add(attack)
        self.to_disk()

    @property
    def attacks_to_run(self):
        return self._attacks_to_run

    @property
    def clean_accuracy(self):
        return self._clean_accuracy

    @clean_accuracy.setter
    def clean_accuracy(self, accuracy):
        self._clean_accuracy = accuracy


class _Logger:
    """autoattack/other_utils.py Logger: print, and append to a file when a path is given."""

    def __init__(self, log_path=None):
        self.log_path = log_path

    def log(self, s):
        print(s)
        if self.log_path is not None:
            with open(self.log_path, "a") as f:
                f.write(s + "\n")
                f.flush()


class AutoAttack:
    def __init__(self, model, norm="Linf", eps=0.3, seed=None, verbose=True, attacks_to_run=(),
                 version="standard", is_tf_model=False, device="cuda", log_path=None, alpha=None,
                 iterations_apgd=100, use_rs=True):
        assert norm in ("Linf", "L2", "L1")                                        # :18
        if norm == "L1" or is_tf_model:
            raise NotImplementedError("only Linf / L2, PyTorch models (SURVEY.md §8a)")
        self.model, self.norm, self.epsilon, self.seed = model, norm, eps, seed
        self.verbose, self.version, self.device = verbose, version, device
        self.attacks_to_run = list(attacks_to_run)
        self.is_tf_model = False
        self.logger = _Logger(log_path)
        if version in ("standard", "plus", "rand") and self.attacks_to_run:        # :29-30
            raise ValueError("attacks_to_run will be overridden unless you use version='custom'")
        self.apgd = APGDAttack(model, n_restarts=5, n_iter=iterations_apgd, verbose=False,
                               eps=eps, norm=norm, eot_iter=1, rho=0.75, seed=seed, device=device,
                               logger=self.logger, alpha=alpha, use_rs=use_rs)      # :34-36
        self.apgd_targeted = APGDAttack_targeted(model, n_restarts=1, n_iter=iterations_apgd,
                                                 verbose=False, eps=eps, norm=norm, eot_iter=1,
                                                 rho=0.75, seed=seed, device=device,
                                                 logger=self.logger, alpha=alpha,
                                                 use_rs=use_rs)                     # :47-49
        if version in ("standard", "plus", "rand"):
            self.set_version(version)

    def get_logits(self, x):
        return self.model(x)

    def get_seed(self):
        return time.time() if self.seed is None else self.seed                     # :78-79

    def set_version(self, version="standard"):
        """:300-345 (APGD members of each preset)"""
        if version == "standard":
            wanted = ["apgd-ce", "apgd-t", "fab-t", "square"]
            self.apgd.n_restarts = 1
            self.apgd_targeted.n_restarts = 1
            self.apgd_targeted.n_target_classes = 9
        elif version == "plus":
            wanted = ["apgd-ce", "apgd-dlr", "fab", "square", "apgd-t", "fab-t"]
            self.apgd.n_restarts = 5
            self.apgd_targeted.n_restarts = 1
            self.apgd_targeted.n_target_classes = 9
        elif version == "rand":
            wanted = ["apgd-ce", "apgd-dlr"]
            self.apgd.n_restarts = 1
            self.apgd.eot_iter = 20
        else:
            raise ValueError(f"unknown version {version}")
        dropped = [a for a in wanted if a in _NOT_BUILT]
        self.attacks_to_run = [a for a in wanted if a in _APGD_FAMILY]
        self.dropped_attacks = dropped
        if dropped:
            # never silent: robust accuracy from the APGD members alone is an UPPER bound of the
            # reference's `{version}` AutoAttack number
            msg = (f"AutoAttack version '{version}': {', '.join(dropped)} are not built here (APGD "
                   f"family only, SURVEY.md §2); running {', '.join(self.attacks_to_run)} -- the "
                   f"robust accuracy reported is an upper bound of the full '{version}' ensemble")
            warnings.warn(msg, stacklevel=2)
            if self.verbose:
                self.logger.log(msg)

    def _run_one(self, attack, x, y):
        if attack == "apgd-ce":                                                    # :184-188
            self.apgd.loss = "ce"
            self.apgd.seed = self.get_seed()
            return self.apgd.perturb(x, y)
        if attack == "apgd-dlr":                                                   # :190-194
            self.apgd.loss = "dlr"
            self.apgd.seed = self.get_seed()
            return self.apgd.perturb(x, y)
        if attack == "apgd-t":                                                     # :207-210
            self.apgd_targeted.seed = self.get_seed()
            return self.apgd_targeted.perturb(x, y)
        if attack in _NOT_BUILT:
            raise NotImplementedError(f"{attack}: only the APGD family is built (SURVEY.md §2)")
        raise ValueError("Attack not supported")                                   # :220

    def run_standard_evaluation(self, x_orig, y_orig, bs=250, return_labels=False, state_path=None):
        """:82-261 -- clean pass, then every attack on the still-robust points in batches of
        `bs`; returns the adversarial images (and the predictions on them)."""
        state_path = Path(state_path) if state_path is not None else None
        if state_path is not None and state_path.exists():                         # :88-98
            state = EvaluationState.from_disk(state_path)
            if set(self.attacks_to_run) != state.attacks_to_run:
                raise ValueError("The state was created with a different set of attacks "
                                 "to run. You are probably using the wrong state file.")
            if self.verbose:
                self.logger.log("Restored state from {}".format(state_path))
                self.logger.log("Since the state has been restored, **only** the adversarial "
                                "examples from the current run are going to be returned.")
        else:
            state = EvaluationState(set(self.attacks_to_run), path=state_path)     # :99-103
            state.to_disk()
        attacks_to_run = [a for a in self.attacks_to_run if a not in state.run_attacks]   # :105
        if self.verbose:
            self.logger.log("using {} version including {}{}.".format(
                self.version, ", ".join(attacks_to_run),
                " (APGD members only; not run: {})".format(", ".join(self.dropped_attacks))
                if getattr(self, "dropped_attacks", None) else ""))
            if state.run_attacks:
                self.logger.log("{} was/were already run.".format(", ".join(sorted(state.run_attacks))))
        n = x_orig.shape[0]
        with torch.no_grad():
            # (the reference leaves y_adv undefined on resume, :125-155; here it starts from the labels)
            y_adv = y_orig.clone()
            if state.robust_flags is None:
                robust_flags = torch.zeros(n, dtype=torch.bool, device=x_orig.device)
                for s in range(0, n, bs):                                          # :127-138
                    x = x_orig[s:s + bs].clone().to(self.device)
                    y = y_orig[s:s + bs].clone().to(self.device)
                    out = self.get_logits(x).max(dim=1)[1]
                    y_adv[s:s + bs] = out.to(y_adv.device)
                    robust_flags[s:s + bs] = y.eq(out).to(robust_flags.device)
                state.robust_flags = robust_flags
                robust_accuracy = robust_flags.sum().item() / n
                self.robust_accuracy_dict = {"clean": robust_accuracy}
                state.clean_accuracy = robust_accuracy
                if self.verbose:
                    self.logger.log("initial accuracy: {:.2%}".format(robust_accuracy))
            else:                                                                  # :147-153
                robust_flags = state.robust_flags.to(x_orig.device)
                robust_accuracy = robust_flags.sum().item() / n
                self.robust_accuracy_dict = {"clean": state.clean_accuracy}
                if self.verbose:
                    self.logger.log("initial clean accuracy: {:.2%}".format(state.clean_accuracy))
                    self.logger.log("robust accuracy at the time of restoring the state: "
                                    "{:.2%}".format(robust_accuracy))
            x_adv = x_orig.clone().detach()
            startt = time.time()
            for attack in attacks_to_run:                                          # :157
                num_robust = robust_flags.sum().item()
                if num_robust == 0:
                    break
                lin = torch.nonzero(robust_flags, as_tuple=False).flatten()
                n_batches = math.ceil(num_robust / bs)
                for bi in range(n_batches):
                    idc = lin[bi * bs:min((bi + 1) * bs, num_robust)]
                    x = x_orig[idc].clone().to(self.device)
                    y = y_orig[idc].clone().to(self.device)
                    adv_curr = self._run_one(attack, x, y)
                    out = self.get_logits(adv_curr).max(dim=1)[1]
                    false_batch = ~y.eq(out).to(robust_flags.device)               # :222-226
                    nr = idc[false_batch]
                    robust_flags[nr] = False
                    state.robust_flags = robust_flags                              # :225
                    x_adv[nr] = adv_curr[false_batch].detach().to(x_adv.device)
                    y_adv[nr] = out[false_batch].detach().to(y_adv.device)
                    if self.verbose:
                        self.logger.log("{} - {}/{} - {} out of {} successfully perturbed".format(
                            attack, bi + 1, n_batches, int(false_batch.sum()), x.shape[0]))
                robust_accuracy = robust_flags.sum().item() / n
                self.robust_accuracy_dict[attack] = robust_accuracy
                state.add_run_attack(attack)                                       # :236
                if self.verbose:
                    self.logger.log("robust accuracy after {}: {:.2%} (total time {:.1f} s)".format(
                        attack.upper(), robust_accuracy, time.time() - startt))
            state.to_disk(force=True)                                              # :243
            if self.verbose:                                                       # :246-256
                d = (x_adv - x_orig).reshape(n, -1)
                res = d.abs().max(1)[0] if self.norm == "Linf" else (d ** 2).sum(-1).sqrt()
                self.logger.log("max {} perturbation: {:.5f}, nan in tensor: {}, max: {:.5f}, "
                                "min: {:.5f}".format(self.norm, res.max(), (x_adv != x_adv).sum(),
                                                     x_adv.max(), x_adv.min()))
                self.logger.log("robust accuracy: {:.2%}".format(robust_accuracy))
        return (x_adv, y_adv) if return_labels else x_adv

    def clean_accuracy(self, x_orig, y_orig, bs=250):
        """:263-275"""
        acc = 0.0
        with torch.no_grad():
            for s in range(0, x_orig.shape[0], bs):
                x = x_orig[s:s + bs].clone().to(self.device)
                y = y_orig[s:s + bs].clone().to(self.device)
                acc += float((self.get_logits(x).max(1)[1] == y).float().sum())
        if self.verbose:
            print("clean accuracy: {:.2%}".format(acc / x_orig.shape[0]))
        return acc / x_orig.shape[0]

    def run_standard_evaluation_individual(self, x_orig, y_orig, bs=250, return_labels=False):
        """:277-298"""
        attacks, adv = self.attacks_to_run, {}
        verbose, self.verbose = self.verbose, False
        for c in attacks:
            self.attacks_to_run = [c]
            x_adv, y_adv = self.run_standard_evaluation(x_orig, y_orig, bs=bs, return_labels=True)
            adv[c] = (x_adv, y_adv) if return_labels else x_adv
        self.attacks_to_run, self.verbose = attacks, verbose
        return adv
