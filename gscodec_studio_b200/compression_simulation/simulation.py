"""Per-attribute compression-simulation policies, API of the reference's
``gsplat/compression_simulation/simulation.py`` (``CompressionSimulation`` :14-349,
``STGCompressionSimulation`` :508-780): which attributes are fake-quantised, their bounds and bit
widths, and from which step the entropy model contributes a rate estimate.

The reference spells the policy out as one method per attribute; here it is a table
(`_POLICY_*`) walked by one routine, and every quantise / entropy evaluation is a single fused
kernel (``ops.fake_quantize_ste``, ``entropy_model``).  Observable behaviour kept:
  * ``simulate_compression(splats, step) -> (new_splats, esti_bits_dict)`` with the same keys;
  * untouched attributes are returned as new autograd nodes with identical values (`+ 0.`, :228);
  * the static trainer's policy always quantises in the default "noise" mode (the reference never
    forwards a q_type); the dynamic (STG) policy forwards its
    ``quantization_sim_type`` and therefore clamps the parameters in place in "round" mode;
  * attributes, bounds, bit widths and entropy gates: simulation.py:26-68 and :525-583.
Additive: ``fuse_activations=True`` makes scales / opacities come back already passed through
exp / sigmoid under the keys "scales_act" / "opacities_act" (one pass instead of three); callers
that do not ask for it see the reference's dictionary.
Additive: ``fuse_into_render=True`` (SURVEY.md section 8b's fusion hook) hands back every attribute that is only
fake-quantised in "noise" mode (no rate term this step) as a lazy ``ops.QuantOnLoad`` under the reference's key: no
kernel runs here, ``torch.exp`` / ``torch.sigmoid`` of it stay lazy, and ``rasterization()`` applies quantisation and
activation inside the projection / colour kernels where the raw parameter is loaded.  Any other consumer of the value
gets the materialised tensor (identical values).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Literal, Optional, Tuple

import torch
from torch import Tensor

from .entropy_model import Entropy_factorized_optimized_refactor, Entropy_gaussian
from .ops import QuantOnLoad, fake_quantize_ste, fake_quantize_ste_multi


@dataclass(frozen=True)
class _AttrPolicy:
    bounds: Optional[Tuple[float, float]]  # None: attribute passes through
    bits: Optional[int]
    entropy: bool = False
    entropy_filters: Tuple[int, ...] = (3, 3, 3)
    entropy_view: Optional[str] = None  # how the [N,...] tensor is presented as [N, ch] to the entropy model


_POLICY_STATIC: Dict[str, _AttrPolicy] = {
    "means": _AttrPolicy(None, None),
    "scales": _AttrPolicy((-10, 2), 8, True, (3, 3)),
    "quats": _AttrPolicy((-1, 1), 8, True, (3, 3, 3)),
    "opacities": _AttrPolicy((-15, 15), 8, False, (3, 3, 3), "unsqueeze"),
    "sh0": _AttrPolicy((-2, 4), 8, True, (3, 3), "squeeze"),
    "shN": _AttrPolicy(None, None),
}

_POLICY_STG: Dict[str, _AttrPolicy] = {
    "means": _AttrPolicy(None, None),
    "scales": _AttrPolicy((-10, 2), 8, True, (3, 3, 3)),
    "quats": _AttrPolicy((-1, 1), 8, True, (3, 3, 3)),
    "opacities": _AttrPolicy((-7, 7), 8, False),
    "trbf_center": _AttrPolicy(None, None),
    "trbf_scale": _AttrPolicy(None, None),
    "motion": _AttrPolicy(None, None),
    "omega": _AttrPolicy(None, None),
    "colors": _AttrPolicy((-7.5, 7.5), 8, True, (3, 3)),
    "features_dir": _AttrPolicy((-10, 10), 8, True, (3, 3)),
    "features_time": _AttrPolicy((-10, 10), 8, True, (3, 3)),
}

_FUSED_ACT = {"scales": "exp", "opacities": "sigmoid"}


class _SimulationBase:
    policy: Dict[str, _AttrPolicy]

    def _setup(self, entropy_model_enable, entropy_steps, device, entropy_model_type="factorized_model",
               gate_negative_steps=True):
        self.entropy_model_enable = entropy_model_enable
        self.entropy_steps = entropy_steps
        self.device = device
        self.simulation_option = {k: p.bounds is not None for k, p in self.policy.items()}
        self.q_bitwidth = {k: p.bits for k, p in self.policy.items()}
        self.bds = {k: (list(p.bounds) if p.bounds is not None else None) for k, p in self.policy.items()}
        self.entropy_model_option = {k: p.entropy for k, p in self.policy.items()}
        if entropy_steps is not None and gate_negative_steps:
            # static policy only (simulation.py:70-73): a negative step switches the rate term of that attribute off;
            # a missing key is an error there too.  The dynamic policy has no such gate (simulation.py:575-583):
            # `step > -1` simply holds from step 0
            for name in self.entropy_model_option:
                if entropy_steps[name] < 0:
                    self.entropy_model_option[name] = False
        if not entropy_model_enable:
            return
        if entropy_model_type not in ("factorized_model", "gaussian_model"):
            raise NotImplementedError("Not implemented entropy model type")
        self.entropy_model_type = entropy_model_type
        self.entropy_models = {}
        self.entropy_model_optimizers = {}
        self.entropy_model_schedulers = {}
        for name, p in self.policy.items():
            model, opt, sch = None, None, None
            if p.entropy:
                ch = {"quats": 4}.get(name, 3)
                if entropy_model_type == "gaussian_model":
                    # hash-grid conditioned Gaussian model (simulation.py:99-107, optimizer :124-128, scheduler :142-145)
                    model = Entropy_gaussian(channel=ch).to(device)
                    opt = torch.optim.Adam([
                        {"params": model.param_regressor.hash_grid.parameters(), "lr": 5e-3, "name": "hash_grid"},
                        {"params": model.param_regressor.mlp_regressor.parameters(), "lr": 5e-3, "name": "mlp_regressor"}])
                    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.01 ** (1.0 / (30000 - entropy_steps[name])))
                else:
                    model = Entropy_factorized_optimized_refactor(channel=ch, filters=p.entropy_filters).to(device)
                    opt = torch.optim.Adam([{"params": prm, "lr": 1e-4, "name": n} for n, prm in model.named_parameters()])
            self.entropy_models[name] = model
            self.entropy_model_optimizers[name] = opt
            self.entropy_model_schedulers[name] = sch
        positive = [k for k, v in entropy_steps.items() if v > 0]
        self.entropy_min_step = entropy_steps[min(positive, key=lambda k: entropy_steps[k])] if positive else 0

    # ---- position sampling of the hash-grid model (simulation.py:185-204; the trainer calls _estiblish_bbox, sic)
    def _estiblish_bbox(self, pos: Tensor, k: float = 1.5):
        self.bbox_lower_bound, self.bbox_upper_bound = torch.quantile(pos, torch.tensor([0.01, 0.99], device=pos.device), dim=0)

    def _get_pts_inside_bbox(self, pos: Tensor) -> Tensor:
        return torch.all((pos >= self.bbox_lower_bound) & (pos <= self.bbox_upper_bound), dim=1)

    def _random_sample_pts(self, mask: Tensor, ratio: float = 0.05) -> Tensor:
        random_values = torch.rand_like(mask.float())
        random_values[~mask] = 0
        threshold = torch.quantile(random_values[mask], 1 - ratio)
        return random_values >= threshold

    def _setup_mask(self, ada_mask_opt, ada_mask_step, strategy, device, kwargs):
        self.shN_qat = False
        self.shN_ada_mask_opt = ada_mask_opt
        self.shN_ada_mask_step = ada_mask_step
        self.shN_ada_mask_strategy = strategy
        if not ada_mask_opt:
            return
        if strategy == "learnable":
            from .ada_mask import AnnealingMask

            cap_max = kwargs.get("cap_max", 1_000_000)
            self.shN_ada_mask = AnnealingMask(input_shape=[cap_max, 1, 1], device=device,
                                              annealing_start_iter=ada_mask_step)
            self.shN_ada_mask_optimizer = torch.optim.Adam([{"params": self.shN_ada_mask.parameters(), "lr": 0.01}])
        elif strategy == "gradient":
            pass
        elif strategy is None:
            raise ValueError("'shN_ada_mask_strategy' should not be None")
        else:
            raise NotImplementedError(f"'shN_ada_mask_strategy': {strategy} has not been implemented.")

    # ------------------------------------------------------------------ the one routine
    def _plan_attr(self, name: str, param: Tensor, step: int, q_type: Optional[str], noise: Optional[Tensor], fuse_act: bool):
        """Arguments of this attribute's fake_quantize_ste call (+ whether a rate term follows, the fused activation)."""
        p = self.policy[name]
        lo, hi = p.bounds
        bits = 8 if step < 10_000 else p.bits  # simulation.py:242-245 (both branches give 8 today)
        want_entropy = (self.entropy_model_enable and self.entropy_model_option.get(name, False)
                        and step > self.entropy_steps[name])
        act = _FUSED_ACT.get(name) if (fuse_act and not want_entropy) else None
        item = dict(input=param, lower_bd=lo, upper_bd=hi, bitwidth=bits, noise=noise, activation=act)
        if q_type is not None:
            item["q_type"] = q_type
        return item, want_entropy, act

    def _finish_attr(self, name: str, fq: Dict, want_entropy: bool, choose_idx: Optional[Tensor], pos: Optional[Tensor]):
        p = self.policy[name]
        value = fq["output_value"]
        esti_bits = None
        if want_entropy:
            flat = value
            if p.entropy_view == "squeeze":
                flat = value.squeeze(1)
            elif p.entropy_view == "unsqueeze":
                flat = value.unsqueeze(1)
            if choose_idx is not None:  # hash-grid model: rate of a ~5 % random sample, conditioned on the positions
                esti_bits = self.entropy_models[name](flat[choose_idx], fq["q_step"], pos=pos)
            else:
                esti_bits = self.entropy_models[name](flat, fq["q_step"])
        return value, esti_bits

    def _simulate(self, splats: Dict[str, Tensor], step: int, q_type, noise, fuse_activations, fuse_into_render=False):
        new_splats, esti_bits_dict = {}, {}
        choose_idx, pos = None, None
        if (self.entropy_model_enable and getattr(self, "entropy_model_type", "factorized_model") == "gaussian_model"
                and step > self.entropy_min_step):
            choose_idx = self._random_sample_pts(self._get_pts_inside_bbox(splats["means"]))  # simulation.py:216-220
            pos = splats["means"][choose_idx]
            self._last_sample_idx = choose_idx  # which points carried the rate estimate (debugging / tests)
        # every quantised attribute of the step goes through ONE fused kernel launch (and one more in the backward)
        plans = {name: self._plan_attr(name, splats[name], step, q_type, None if noise is None else noise.get(name),
                                       fuse_activations)
                 for name in splats.keys() if self.policy[name].bounds is not None and self.simulation_option[name]}
        lazy = {}
        if fuse_into_render and q_type in (None, "noise"):
            # quantise on load: nothing is computed here for attributes without a rate term this step
            for name, (item, want_entropy, _act) in plans.items():
                if not want_entropy and splats[name].is_cuda and splats[name].dtype == torch.float32:
                    lazy[name] = QuantOnLoad(splats[name], item["lower_bd"], item["upper_bd"], item["bitwidth"], item["noise"])
        eager = {name: pl for name, pl in plans.items() if name not in lazy}
        if len(eager) > 1:
            fqs = dict(zip(eager, fake_quantize_ste_multi([pl[0] for pl in eager.values()])))
        else:
            fqs = {name: fake_quantize_ste(**pl[0]) for name, pl in eager.items()}
        for name in splats.keys():
            esti_bits_dict[name] = None
            if name in lazy:
                act = _FUSED_ACT.get(name) if fuse_activations else None
                new_splats[name + "_act" if act is not None else name] = lazy[name] if act is None else lazy[name]._activated(act)
            elif name in plans:
                _item, want_entropy, act = plans[name]
                value, bits = self._finish_attr(name, fqs[name], want_entropy, choose_idx, pos)
                new_splats[name + "_act" if act is not None else name] = value
                esti_bits_dict[name] = bits
            elif name in self.returns_self:
                # the reference routes these through a routine that hands back the tensor itself,
                # optionally times the annealed mask (simulation.py:318-323)
                value = splats[name]
                if self.shN_ada_mask_opt and self.shN_ada_mask_strategy == "learnable" and step > self.shN_ada_mask_step:
                    value = self.shN_ada_mask(value, step)
                new_splats[name] = value
            else:
                new_splats[name] = splats[name] + 0.0  # new autograd node, same values (simulation.py:228)
        return new_splats, esti_bits_dict


class CompressionSimulation(_SimulationBase):
    """Static-scene policy (means / scales / quats / opacities / sh0 / shN)."""

    policy = _POLICY_STATIC
    returns_self = frozenset({"shN"})

    def __init__(self, entropy_model_enable: bool = False,
                 entropy_model_type: Literal["factorized_model", "gaussian_model"] = "factorized_model",
                 entropy_steps: Dict[str, int] = None, device=None, ada_mask_opt: bool = False,
                 ada_mask_step: int = 10_000, ada_mask_strategy: Optional[str] = "learnable", **kwargs) -> None:
        self.entropy_model_type = entropy_model_type
        if entropy_steps is None:
            entropy_steps = {k: -1 for k in self.policy}
        self._setup(entropy_model_enable, entropy_steps, device, entropy_model_type)
        self.simulation_option["shN"] = True  # as in the reference: shN is "simulated" by an identity / mask
        self._setup_mask(ada_mask_opt, ada_mask_step, ada_mask_strategy, device, kwargs)

    def simulate_compression(self, splats: Dict[str, Tensor], step: int, noise: Optional[Dict[str, Tensor]] = None,
                             fuse_activations: bool = False, fuse_into_render: bool = False):
        """-> (new_splats, esti_bits_dict).  `noise` (optional, per attribute) replaces the in-kernel draw."""
        return self._simulate(splats, step, None, noise, fuse_activations, fuse_into_render)

    def shN_gradient_threshold(self, param: torch.nn.Parameter, step: int) -> None:
        """Zero the gradients of all-zero shN rows whose gradient norm is small (simulation.py:326-349)."""
        value = param.data
        zero_rows = (value == 0).all(dim=-1).all(dim=-1)
        non_zero_ratio = 1 - zero_rows.sum() / zero_rows.size(0)
        threshold = 2e-3 if non_zero_ratio < 0.10 else 100
        small = torch.norm(param.grad, p=2, dim=(-2, -1)) < threshold
        param.grad[torch.logical_and(zero_rows, small)] = 0


class STGCompressionSimulation(_SimulationBase):
    """Dynamic-scene (Spacetime Gaussian) policy; forwards its quantisation type ("round" in the shipped
    configs, examples/simple_trainer_dyngs.py:1325-1331)."""

    policy = _POLICY_STG
    returns_self = frozenset()

    def __init__(self, quantization_sim_type: Optional[Literal["round", "noise", "vq"]] = None,
                 entropy_model_enable: bool = False, entropy_steps: Dict[str, int] = None, device=None,
                 ada_mask_opt: bool = False, ada_mask_step: int = 10_000, **kwargs) -> None:
        self.quantization_sim_type = quantization_sim_type
        if entropy_steps is None:
            entropy_steps = {k: -1 for k in self.policy}
        self._setup(entropy_model_enable, entropy_steps, device, gate_negative_steps=False)
        self._setup_mask(ada_mask_opt, ada_mask_step, "learnable", device, kwargs)

    def simulate_compression(self, splats: Dict[str, Tensor], step: int, noise: Optional[Dict[str, Tensor]] = None,
                             fuse_activations: bool = False, fuse_into_render: bool = False):
        q_type = self.quantization_sim_type
        if q_type is None:
            raise NotImplementedError("quantization_sim_type must be 'round' or 'noise'")
        return self._simulate(splats, step, q_type, noise, fuse_activations=False, fuse_into_render=fuse_into_render)
