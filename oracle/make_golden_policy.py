"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_policy.npz (+ .json) by running the REFERENCE's own
gsplat/compression_simulation/simulation.py on the CPU of the build container:

    python -m oracle.make_golden_policy            (needs /root/reference; the GPU box only reads the fixture)

The reference package imports the CUDA-only `_gridencoder` extension at import time (gaussian_distribution_model.py:8);
a stub module stands in for it -- nothing on the factorized-model path touches it.  What is recorded:

  * the policy tables of CompressionSimulation and STGCompressionSimulation (simulation.py:26-68, 527-580): which
    attributes are simulated, bit widths, bounds, which get a rate term, and the negative-step gating (:70-73);
  * state_dict keys / shapes of the three entropy models (checkpoint compatibility, simple_trainer.py:1093-1095, 1474);
  * a functional run of CompressionSimulation.simulate_compression (noise mode) and of STGCompressionSimulation
    (round mode) at steps before / between / after the entropy gates on seeded inputs, with the uniform noise the
    reference drew captured call by call, the simulated values and the estimated bits, and the entropy-model weights.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GSPLAT_REF", "/root/reference")
STEPS_STATIC = {"means": -1, "scales": 10, "quats": 10, "opacities": -1, "sh0": 20, "shN": -1}


def _reference():
    sys.modules.setdefault("_gridencoder", types.ModuleType("_gridencoder"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from gsplat.compression_simulation import simulation

    return simulation


class _NoiseTap:
    """Records every tensor.uniform_(-0.5, 0.5) the reference draws (ops.py:46)."""

    def __enter__(self):
        self.draws, self._orig = [], torch.Tensor.uniform_
        tap = self

        def uniform_(t, *a, **k):
            out = tap._orig(t, *a, **k)
            tap.draws.append(out.detach().clone())
            return out

        torch.Tensor.uniform_ = uniform_
        return self

    def __exit__(self, *a):
        torch.Tensor.uniform_ = self._orig


def main():
    S = _reference()
    torch.manual_seed(0)
    np.random.seed(0)
    out, meta = {}, {}

    # ---------------------------------------------------------------- static policy
    sim = S.CompressionSimulation(entropy_model_enable=True, entropy_steps=dict(STEPS_STATIC), device="cpu")
    meta["static"] = {"simulation_option": sim.simulation_option, "q_bitwidth": sim.q_bitwidth, "bds": sim.bds,
                      "entropy_model_option": sim.entropy_model_option, "entropy_min_step": sim.entropy_min_step,
                      "entropy_steps": STEPS_STATIC}
    meta["static"]["state_dict"] = {k: {n: list(v.shape) for n, v in m.state_dict().items()}
                                    for k, m in sim.entropy_models.items() if m is not None}
    meta["static"]["optimizer_param_names"] = {k: [g["name"] for g in o.param_groups]
                                               for k, o in sim.entropy_model_optimizers.items() if o is not None}
    for k, m in sim.entropy_models.items():
        if m is not None:
            with torch.no_grad():  # move the weights off their constant initialisation
                for p in m.parameters():
                    p.add_(torch.randn_like(p) * 0.3)
            for n, v in m.state_dict().items():
                out[f"static/model/{k}/{n}"] = v.detach().numpy()
    g = torch.Generator().manual_seed(11)
    for N in (777, 1024):  # 1024: N % 32 == 0 takes the reference's full extra pad block (entropy_model.py:223)
        splats = {"means": torch.randn(N, 3, generator=g) * 2, "scales": torch.randn(N, 3, generator=g) * 4 - 4,
                  "quats": torch.randn(N, 4, generator=g) * 0.7, "opacities": torch.randn(N, generator=g) * 9,
                  "sh0": torch.randn(N, 1, 3, generator=g) * 2 + 1, "shN": torch.randn(N, 15, 3, generator=g) * 0.1}
        for k, v in splats.items():
            out[f"static/{N}/in/{k}"] = v.numpy().copy()
        for step in (5, 15, 25):
            with _NoiseTap() as tap:
                new, bits = sim.simulate_compression({k: v.clone() for k, v in splats.items()}, step)
            names = [k for k in splats if sim.simulation_option[k] and sim.bds[k] is not None]
            assert len(tap.draws) == len(names), (len(tap.draws), names)
            for k, d in zip(names, tap.draws):
                out[f"static/{N}/{step}/noise/{k}"] = d.numpy()
            for k, v in new.items():
                out[f"static/{N}/{step}/out/{k}"] = v.detach().numpy()
            for k, b in bits.items():
                if b is not None:
                    out[f"static/{N}/{step}/bits/{k}"] = b.detach().numpy()
            meta["static"].setdefault("bits_present", {})[f"{N}/{step}"] = sorted(k for k, b in bits.items() if b is not None)

    # ---------------------------------------------------------------- dynamic (STG) policy
    steps_stg = {"means": -1, "scales": 10, "quats": 10, "opacities": -1, "colors": 10, "features_dir": 10, "features_time": 10}
    stg = S.STGCompressionSimulation(quantization_sim_type="round", entropy_model_enable=True, entropy_steps=dict(steps_stg),
                                     device="cpu")
    meta["stg"] = {"simulation_option": stg.simulation_option, "q_bitwidth": stg.q_bitwidth, "bds": stg.bds,
                   "entropy_model_option": stg.entropy_model_option, "entropy_steps": steps_stg}
    meta["stg"]["state_dict"] = {k: {n: list(v.shape) for n, v in m.state_dict().items()}
                                 for k, m in stg.entropy_models.items() if m is not None}
    for k, m in stg.entropy_models.items():
        if m is not None:
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(torch.randn_like(p) * 0.3)
            for n, v in m.state_dict().items():
                out[f"stg/model/{k}/{n}"] = v.detach().numpy()
    N = 555
    splats = {"means": torch.randn(N, 3, generator=g), "scales": torch.randn(N, 3, generator=g) * 4 - 4,
              "quats": torch.randn(N, 4, generator=g) * 0.7, "opacities": torch.randn(N, generator=g) * 5,
              "trbf_center": torch.rand(N, 1, generator=g), "trbf_scale": torch.randn(N, 1, generator=g),
              "motion": torch.randn(N, 9, generator=g) * 0.1, "omega": torch.randn(N, 4, generator=g) * 0.1,
              "colors": torch.randn(N, 3, generator=g) * 5, "features_dir": torch.randn(N, 3, generator=g) * 6,
              "features_time": torch.randn(N, 3, generator=g) * 6}
    for k, v in splats.items():
        out[f"stg/in/{k}"] = v.numpy().copy()
    for step in (5, 15):
        ins = {k: v.clone() for k, v in splats.items()}
        new, bits = stg.simulate_compression(ins, step)
        for k, v in new.items():
            out[f"stg/{step}/out/{k}"] = v.detach().numpy()
        for k, v in ins.items():  # round mode clamps the PARAMETER in place (ops.py:63)
            out[f"stg/{step}/param_after/{k}"] = v.detach().numpy()
        for k, b in bits.items():
            if b is not None:
                out[f"stg/{step}/bits/{k}"] = b.detach().numpy()
        meta["stg"].setdefault("bits_present", {})[str(step)] = sorted(k for k, b in bits.items() if b is not None)

    dst = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(dst, "ref_policy.npz"), **out)
    with open(os.path.join(dst, "ref_policy.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", {k: meta[k]["bits_present"] for k in meta})


if __name__ == "__main__":
    main()
