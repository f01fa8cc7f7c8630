"""Host glue of the planning / training loop around the hot path: the parts of ``ActiveNeRFMapper``
(scripts/pipeline.py) that call the renderer, the scorer and the training step, with the reference's
on-disk formats.  The simulator, the trajectory sampler, plotting and image writing are out of scope
(SURVEY.md section 2): trajectories and training batches come in as arrays / callables.

  ActiveNeRFMapper.probablistic_uncertainty      scripts/pipeline.py:666-798   (one trajectory -> score + log entry)
  planning(): loop over the sampled trajectories scripts/pipeline.py:1077-1085 (folded into ONE batched scorer call)
  ActiveNeRFMapper.nerf_training                 scripts/pipeline.py:354-664   (ensemble training loop, checkpoints)
  ActiveNeRFMapper.render                        scripts/pipeline.py:918-1023  (the model-side renders and their
                                                                                8-bit images; no simulator / cv2)
  checkpoint + uncertainty.npy                   scripts/pipeline.py:630-634, 1256-1274
"""
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .data_proc import Dataset
from .scoring import PredictiveInformationScorer, trajector_uncertainty
from .training import EnsembleTrainer


class ActiveNeRFMapper:
    """The render / score / train members of the reference class of the same name.  ``config`` uses the reference's
    keys (scripts/config_*.yaml): img_w, img_h, hfov, near_plane, render_step_size, cone_angle, alpha_thre,
    planning_step, num_traj, cuda."""

    def __init__(self, radiance_fields: Sequence[torch.nn.Module], estimators: Sequence[torch.nn.Module],
                 optimizers: Sequence[torch.optim.Optimizer], config: Dict, schedulers=None, process_group=None):
        self.radiance_fields, self.estimators = list(radiance_fields), list(estimators)
        self.optimizers = list(optimizers)
        self.schedulers = list(schedulers) if schedulers is not None else [None] * len(self.optimizers)
        self.config_file = dict(config)
        c = self.config_file
        self.focal = 0.5 * c["img_w"] / np.tan(c["hfov"] / 2)  # pipeline.py:203-205
        self.device = torch.device(c.get("cuda", "cuda:0"))
        self.trajector_uncertainty_list: List[list] = [[] for _ in range(c.get("planning_step", 1))]  # :136-138
        self.process_group = process_group
        self._scorer = None
        self._trainer = None

    # ---- scoring -------------------------------------------------------------------------------------------
    def _render_opts(self):
        c = self.config_file
        return dict(near_plane=c["near_plane"], render_step_size=c["render_step_size"], cone_angle=c["cone_angle"],
                    alpha_thre=c["alpha_thre"])

    def scorer(self, scale: float = 0.1) -> PredictiveInformationScorer:
        c = self.config_file
        key = (scale, tuple(id(f) for f in self.radiance_fields), tuple(id(e) for e in self.estimators),
               tuple(sorted(self._render_opts().items())), c["img_w"], c["img_h"])
        if self._scorer is None or self._scorer[0] != key:
            s = PredictiveInformationScorer(self.radiance_fields, self.estimators, c["img_w"], c["img_h"], self.focal,
                                            scale=scale, device=self.device, views_per_batch=c.get("views_per_batch"),
                                            **self._render_opts())
            self._scorer = (key, s)
        return self._scorer[1]

    def score_trajectories(self, trajectories: Sequence[np.ndarray], step: int, scale: float = 0.1):
        """The planner's ``for i in range(num_traj): uncertainty = self.probablistic_uncertainty(traj[i], step)``
        (pipeline.py:1077-1085) as one batched, view-sharded call.  Returns (uncertainties [n_traj], best_index) and
        appends one ``[rgb, depth, 3 * sem, 2 * occ]`` entry per trajectory to ``trajector_uncertainty_list[step - 1]``
        exactly as the reference does (:783-790)."""
        terms = self.scorer(scale).score_trajectories([np.asarray(t) for t in trajectories],
                                                      process_group=self.process_group)
        for t in terms:
            self.trajector_uncertainty_list[step - 1].append(t.tolist())
        unc = terms.sum(1)
        return unc, int(np.argmax(unc))

    def probablistic_uncertainty(self, trajectory, step) -> float:
        """Drop-in for ActiveNeRFMapper.probablistic_uncertainty(trajectory, step) -> float (pipeline.py:666-798)."""
        return float(self.score_trajectories([trajectory], step)[0][0])

    def trajector_uncertainty(self, trajectory, step):
        """Legacy scorer of the "random" policy (pipeline.py:800-916)."""
        c = self.config_file
        log = []
        out = trajector_uncertainty(self.radiance_fields, self.estimators, trajectory, step, img_w=c["img_w"],
                                    img_h=c["img_h"], focal=self.focal, scale=0.1, device=self.device, log=log,
                                    **self._render_opts())
        self.trajector_uncertainty_list[step - 1].append(log[0])
        return out

    # ---- training ------------------------------------------------------------------------------------------
    def nerf_training(self, steps: int, fetch_batch: Callable[[int], Dict], planning_step: int = -1,
                      checkpoint_dir: Optional[str] = None, checkpoint_every: int = 1000):
        """The ensemble loop of nerf_training (pipeline.py:403-532): per step, every member draws a batch
        (``fetch_batch(model_idx)``), updates its occupancy grid, renders in train mode, takes the loss and an Adam step.
        ``occ_thre`` follows the reference's schedule by planning step (:452-474); every ``checkpoint_every`` steps
        member 0 is checkpointed (:616-636)."""
        occ_thre = 1e-3 if planning_step == -1 else 1e-2 if planning_step == -10 else 1e-3 if planning_step < 5 else 3e-3
        if self._trainer is None or self._trainer.opts["occ_thre"] != occ_thre:
            self._trainer = EnsembleTrainer(self.radiance_fields, self.estimators, self.optimizers, occ_thre=occ_thre,
                                            schedulers=self.schedulers, process_group=self.process_group,
                                            **self._render_opts())
        tr = self._trainer
        idx = [0]

        def fetch():
            b = fetch_batch(idx[0] % len(self.radiance_fields))
            idx[0] += 1
            return b

        for step in range(steps):
            tr.step(fetch, step)
            if checkpoint_dir is not None and (step + 1) % checkpoint_every == 0:
                self.save_checkpoint(0, os.path.join(checkpoint_dir, f"model_step{step + 1}.pth"))
        return tr

    def planning_round(self, trajectories, step: int, training_steps: int, fetch_batch, scale: float = 0.1):
        """One iteration of planning() without the simulator: score the candidates, pick the best (:1085), retrain
        (:1211).  Returns (best_index, uncertainties)."""
        for m in self.radiance_fields + self.estimators:
            m.eval()
        unc, best = self.score_trajectories(trajectories, step, scale)
        self.nerf_training(training_steps, fetch_batch, planning_step=step)
        return best, unc

    # ---- visualisation renders -----------------------------------------------------------------------------
    def render(self, traj) -> Dict[str, np.ndarray]:
        """The model side of ActiveNeRFMapper.render (pipeline.py:955-1021): member 0's full-resolution renders of the
        poses (``Dataset.render_image_from_pose`` with scale = 1, downsample = 1) and the 8-bit images the reference
        writes -- ``pd_rgb = float32(rgb * 255)``, ``pd_dep = clip(depth * 25, 0, 255)``, ``pd_occ = clip(acc * 255,
        0, 255)``, ``pd_sem = argmax(sem logits)`` (the reference maps the label through habitat's colour table).
        Returns the float64 predictions and the images; writing files (cv2) stays with the caller."""
        c = self.config_file
        out = Dataset.render_image_from_pose(self.radiance_fields[0], self.estimators[0], np.asarray(traj), c["img_w"],
                                             c["img_h"], self.focal, c["near_plane"], c["render_step_size"], 1,
                                             c["cone_angle"], c["alpha_thre"], 1, self.device)
        rgb, dep, acc = out[0], out[1], out[2]
        res = {"rgb_predictions": rgb, "depth_predictions": dep, "acc_predictions": acc,
               "pd_rgb": np.float32(rgb * 255), "pd_dep": np.clip(dep * 25, 0, 255), "pd_occ": np.clip(acc * 255, 0, 255)}
        if len(out) > 3:
            res["sem_predictions"] = out[3]
            res["pd_sem"] = np.argmax(out[3], axis=-1)
        return res

    # ---- on-disk formats -----------------------------------------------------------------------------------
    def save_checkpoint(self, i: int, path: str) -> None:
        """{"occ_grid": estimator.binaries, "model": state_dict, "optimizer_state_dict": ...} (pipeline.py:630-634,
        1269-1274).  The model keys are the reference's (``mlp_base.params`` ... flat fp32, tcnn parameter order)."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save({"occ_grid": self.estimators[i].binaries, "model": self.radiance_fields[i].state_dict(),
                    "optimizer_state_dict": self.optimizers[i].state_dict()}, path)

    def load_checkpoint(self, i: int, path: str) -> None:
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.radiance_fields[i].load_state_dict(ck["model"])
        self.estimators[i].binaries = ck["occ_grid"].to(self.estimators[i].binaries.device)
        self.optimizers[i].load_state_dict(ck["optimizer_state_dict"])

    def save_all(self, save_path: str) -> None:
        """End of pipeline(): ``uncertainty.npy`` (:1256-1257) and one checkpoint per member (:1262-1274)."""
        os.makedirs(os.path.join(save_path, "checkpoints"), exist_ok=True)
        np.save(os.path.join(save_path, "uncertainty.npy"), np.array(self.trajector_uncertainty_list))
        for i in range(len(self.radiance_fields)):
            self.save_checkpoint(i, os.path.join(save_path, "checkpoints", f"model_{i}.pth"))
