"""Densify / prune mechanics of the Gaussian parameter store (second half).

Mirrors GaussianModel.{replace_tensor_to_optimizer, _prune_optimizer, prune_points, cat_tensors_to_optimizer,
densification_postfix, densify_and_split, densify_and_clone, densify_and_prune, reset_opacity} of taekkii/deblurgs
(scene/gaussian_model.py:247-254, 300-454) on `GaussianParams` + `FusedAdam` (or any optimizer with torch's
`param_groups` / `state` layout): same selection rules, same new-Gaussian construction, same optimizer-state
surgery (moments masked / zero-extended, the per-tensor step count kept), 'curve_*' groups left alone.

These are bulk gather / concatenate operations executed every `densification_interval` (200) iterations, not per
step: they stay torch ops on whatever device the parameters live on (so they are testable on CPU); the per-step
part -- accumulating the statistics they consume -- is fused into the backward kernel (`DensificationStats`).
The WHEN (thresholds, schedules) remains the caller's policy, as in the reference's train.py:188-199.
"""
import torch
import torch.nn as nn


def build_rotation(r):
    """Quaternions (r, x, y, z), not necessarily normalised, -> rotation matrices (utils/general_utils.py:117-138)."""
    q = r / torch.sqrt((r * r).sum(dim=1))[:, None]
    w, x, y, z = q.unbind(1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


_GROUP_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
               "scaling": "_scaling", "rotation": "_rotation"}


class DensificationMixin:
    """Needs: the six raw parameter attributes, `optimizer` (param_groups named as in training_setup),
    `get_scaling` / `get_opacity` / `get_xyz`, `scale_lower_bound`, and the statistics buffers."""

    percent_dense = 0.01          # arguments/__init__.py:94
    alpha_lower_bound = 0.0       # arguments/__init__.py:68

    # ---- optimizer-state surgery -----------------------------------------------------------------
    def _swap_param(self, group, new_tensor, new_state_fn):
        old = group["params"][0]
        stored = self.optimizer.state.get(old, None)
        new = nn.Parameter(new_tensor.requires_grad_(True))
        if stored is not None:
            stored["exp_avg"] = new_state_fn(stored["exp_avg"])
            stored["exp_avg_sq"] = new_state_fn(stored["exp_avg_sq"])
            del self.optimizer.state[old]
            self.optimizer.state[new] = stored
        group["params"][0] = new
        return new

    def _assign(self, tensors):
        for name, attr in _GROUP_ATTR.items():
            if name in tensors:
                setattr(self, attr, tensors[name])

    def replace_tensor_to_optimizer(self, tensor, name):
        out = {}
        for group in self.optimizer.param_groups:
            if group["name"] == name:
                out[name] = self._swap_param(group, tensor, lambda m: torch.zeros_like(tensor))
        return out

    def _prune_optimizer(self, mask):
        out = {}
        for group in self.optimizer.param_groups:
            if "curve_" in group["name"]:
                continue
            out[group["name"]] = self._swap_param(group, group["params"][0].detach()[mask], lambda m: m[mask])
        return out

    def cat_tensors_to_optimizer(self, tensors_dict):
        out = {}
        for group in self.optimizer.param_groups:
            assert len(group["params"]) == 1
            if group["name"] not in tensors_dict:
                continue
            ext = tensors_dict[group["name"]]
            out[group["name"]] = self._swap_param(
                group, torch.cat((group["params"][0].detach(), ext), dim=0),
                lambda m, ext=ext: torch.cat((m, torch.zeros_like(ext)), dim=0))
        return out

    # ---- store-level operations ------------------------------------------------------------------
    @torch.no_grad()
    def prune_points(self, mask):
        """Remove the Gaussians where `mask` is True; parameters, Adam moments and statistics are compacted."""
        keep = ~mask
        self._assign(self._prune_optimizer(keep))
        self._ensure_stats()
        self.xyz_gradient_accum = self.xyz_gradient_accum[keep]
        self.denom = self.denom[keep]
        self.max_radii2D = self.max_radii2D[keep]

    @torch.no_grad()
    def densification_postfix(self, new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling,
                              new_rotation):
        """Append new Gaussians (zero Adam moments) and reset the densification statistics."""
        self._assign(self.cat_tensors_to_optimizer({"xyz": new_xyz, "f_dc": new_features_dc,
                                                    "f_rest": new_features_rest, "opacity": new_opacities,
                                                    "scaling": new_scaling, "rotation": new_rotation}))
        P, dev = self._xyz.shape[0], self._xyz.device
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.max_radii2D = torch.zeros(P, device=dev)

    def _scaling_inverse(self, s):
        # LowerBoundLog (scene/gaussian_activation.py:54-64)
        return torch.log((s - self.scale_lower_bound).clamp_min(0.001))

    @torch.no_grad()
    def densify_and_split(self, grads, grad_threshold, scene_extent, N=2, generator=None):
        """Large Gaussians with a large view-space gradient are replaced by N samples of themselves, scaled down
        by 0.8 N (scene/gaussian_model.py:396-420)."""
        P, dev = self._xyz.shape[0], self._xyz.device
        padded = torch.zeros(P, device=dev)
        padded[:grads.shape[0]] = grads.squeeze()
        scaling = self.get_scaling
        sel = (padded >= grad_threshold) & (scaling.max(dim=1).values > self.percent_dense * scene_extent)
        stds = scaling[sel].repeat(N, 1)
        samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator)
        rots = build_rotation(self._rotation[sel]).repeat(N, 1, 1)
        new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + self._xyz[sel].repeat(N, 1)
        new_scaling = self._scaling_inverse(scaling[sel].repeat(N, 1) / (0.8 * N))
        self.densification_postfix(new_xyz, self._features_dc[sel].repeat(N, 1, 1),
                                   self._features_rest[sel].repeat(N, 1, 1), self._opacity[sel].repeat(N, 1),
                                   new_scaling, self._rotation[sel].repeat(N, 1))
        self.prune_points(torch.cat((sel, torch.zeros(N * int(sel.sum()), device=dev, dtype=torch.bool))))

    @torch.no_grad()
    def densify_and_clone(self, grads, grad_threshold, scene_extent):
        """Small Gaussians with a large view-space gradient are duplicated in place (scene/gaussian_model.py:422-436)."""
        sel = (torch.norm(grads, dim=-1) >= grad_threshold) & \
              (self.get_scaling.max(dim=1).values <= self.percent_dense * scene_extent)
        self.densification_postfix(self._xyz[sel], self._features_dc[sel], self._features_rest[sel],
                                   self._opacity[sel], self._scaling[sel], self._rotation[sel])

    @torch.no_grad()
    def densify_and_prune(self, max_grad, extent, generator=None):
        """One densification round from the accumulated statistics (scene/gaussian_model.py:438-449)."""
        self._ensure_stats()
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, generator=generator)
        min_opacity = self.alpha_lower_bound + (1 - self.alpha_lower_bound) * 0.005
        self.prune_points((self.get_opacity < min_opacity).squeeze(-1))

    @torch.no_grad()
    def reset_opacity(self, new_opacity=None):
        """Cap every opacity at `new_opacity` (default 0.1) and zero its Adam moments (scene/gaussian_model.py:247-254)."""
        if new_opacity is None:
            lb = self.alpha_lower_bound
            new_opacity = lb + (1 - lb) * 0.1
        opac = self.get_opacity
        capped = torch.min(opac, torch.ones_like(opac) * new_opacity).clamp(0.0, 1.0)
        self._assign(self.replace_tensor_to_optimizer(capped, "opacity"))
