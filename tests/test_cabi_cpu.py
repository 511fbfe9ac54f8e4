"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/gsb200.h declares; the Python operator layer mirrors the reference's names; nothing silently
falls back to the CPU.  No compute calls (there is no GPU here)."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gscodec_studio_b200 import _build, _cabi

    _build.build_library()
    lib = _cabi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gsb200.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_cabi.SIGNATURES) == declared
    assert lib.gsb200_abi_version() == _cabi.ABI_VERSION
    assert lib.gsb200_grad_record(3) == 12 and lib.gsb200_grad_record(32) == 40


def test_library_is_sm100a_only():
    import subprocess
    from gscodec_studio_b200 import _cabi

    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_operator_layer_mirrors_reference_names():
    import gscodec_studio_b200 as gs
    from gscodec_studio_b200.cuda import _wrapper

    expect = {
        "spherical_harmonics": ["degrees_to_use", "dirs", "coeffs", "masks"],
        "fully_fused_projection": ["means", "covars", "quats", "scales", "viewmats", "Ks", "width", "height", "eps2d",
                                   "near_plane", "far_plane", "radius_clip", "packed", "sparse_grad",
                                   "calc_compensations", "camera_model"],
        "isect_tiles": ["means2d", "radii", "depths", "tile_size", "tile_width", "tile_height", "sort", "packed",
                        "n_cameras", "camera_ids", "gaussian_ids"],
        "isect_offset_encode": ["isect_ids", "n_cameras", "tile_width", "tile_height"],
        "rasterize_to_pixels": ["means2d", "conics", "colors", "opacities", "image_width", "image_height", "tile_size",
                                "isect_offsets", "flatten_ids", "backgrounds", "masks", "packed", "absgrad"],
    }
    for fn, params in expect.items():
        assert list(inspect.signature(getattr(_wrapper, fn)).parameters) == params, fn
        assert getattr(gs, fn) is getattr(_wrapper, fn)
    sig = inspect.signature(gs.rasterization)
    ref_order = ["means", "quats", "scales", "opacities", "colors", "viewmats", "Ks", "width", "height", "near_plane",
                 "far_plane", "radius_clip", "eps2d", "sh_degree", "packed", "tile_size", "backgrounds", "render_mode",
                 "sparse_grad", "absgrad", "rasterize_mode", "channel_chunk", "distributed", "camera_model", "covars"]
    assert list(sig.parameters)[: len(ref_order)] == ref_order
    assert sig.parameters["packed"].default is True and sig.parameters["tile_size"].default == 16
    assert sig.parameters["eps2d"].default == 0.3 and sig.parameters["far_plane"].default == 1e10


def test_no_cpu_fallback():
    import gscodec_studio_b200 as gs

    with pytest.raises(RuntimeError, match="CUDA"):
        gs.spherical_harmonics(1, torch.zeros(4, 3), torch.zeros(4, 4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        gs.fully_fused_projection(torch.zeros(4, 3), None, torch.ones(4, 4), torch.ones(4, 3), torch.eye(4)[None],
                                  torch.eye(3)[None], 16, 16)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gscodec_studio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "_helper.py", os.path.join(dirpath, f)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) works without a GPU and prints one
    JSON line with the keys of the contract; one bounded step of the headline workload on the host cores."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--scene-grid", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["metric"].startswith("Mpix/s fwd+bwd")


def test_profile_helpers_are_inert_without_timeit(monkeypatch):
    """gscodec_studio_b200.profile mirrors gsplat/profile.py's surface; without TIMEIT=1 it must cost nothing and
    record nothing (and never touch CUDA)."""
    from gscodec_studio_b200 import profile

    monkeypatch.delenv("TIMEIT", raising=False)
    profile.profiler.clear()

    @profile.timeit()
    def f(x):
        return x + 1

    with profile.timeit(name="region"):
        assert f(1) == 2
    assert profile.profiler_summary() == {} and f.__name__ == "f"


def test_compression_simulation_host_logic_with_stubbed_kernels(monkeypatch):
    """Host-side policy of CompressionSimulation on the CPU with the two kernels stubbed out: which attributes are
    quantised, which get a rate estimate and from which step, pass-through of means / shN -- for the default
    factorized model and for the (opt-in, experimental) hash-grid model with its position sampling."""
    from gscodec_studio_b200.compression_simulation import simulation as S

    calls = []

    def fake_fq(x, lo, hi, bits, noise=None, activation=None, **kw):
        calls.append((tuple(x.shape), lo, hi, bits, activation))
        return {"output_value": torch.clamp(x, lo, hi), "q_step": (hi - lo) / (2 ** bits - 1)}

    class StubModel(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))
            self.param_regressor = torch.nn.ModuleDict({"hash_grid": torch.nn.Linear(1, 1), "mlp_regressor": torch.nn.Linear(1, 1)})

        def forward(self, x, q, pos=None):
            assert pos is None or pos.shape[0] == x.shape[0]
            return x.abs() * self.w

    monkeypatch.setattr(S, "fake_quantize_ste", fake_fq)
    monkeypatch.setattr(S, "fake_quantize_ste_multi", lambda items: [
        fake_fq(it["input"], it["lower_bd"], it["upper_bd"], it["bitwidth"], noise=it.get("noise"), activation=it.get("activation"))
        for it in items])
    monkeypatch.setattr(S, "Entropy_factorized_optimized_refactor", StubModel)
    monkeypatch.setattr(S, "Entropy_gaussian", StubModel)
    g = torch.Generator().manual_seed(0)
    N = 400
    splats = {"means": torch.randn(N, 3, generator=g), "scales": torch.randn(N, 3, generator=g), "quats": torch.randn(N, 4, generator=g),
              "opacities": torch.randn(N, generator=g), "sh0": torch.randn(N, 1, 3, generator=g), "shN": torch.randn(N, 15, 3, generator=g)}
    steps = {"means": -1, "scales": 10, "quats": 10, "opacities": -1, "sh0": 20, "shN": -1}
    sim = S.CompressionSimulation(entropy_model_enable=True, entropy_steps=steps, device="cpu")
    new, bits = sim.simulate_compression(splats, 15)
    assert set(new) == set(splats) and new["shN"] is splats["shN"] and torch.equal(new["means"], splats["means"])
    assert bits["scales"].shape == (N, 3) and bits["quats"].shape == (N, 4) and bits["sh0"] is None and bits["opacities"] is None
    assert [c[1:3] for c in calls] == [(-10, 2), (-1, 1), (-15, 15), (-2, 4)]
    new, bits = sim.simulate_compression(splats, 25)
    assert bits["sh0"].shape == (N, 3)
    # hash-grid policy: ~5 % of the points inside the 1-99 % box, rate conditioned on their positions
    gsim = S.CompressionSimulation(entropy_model_enable=True, entropy_model_type="gaussian_model", entropy_steps=steps, device="cpu")
    gsim._estiblish_bbox(splats["means"])
    assert isinstance(gsim.entropy_model_schedulers["scales"], torch.optim.lr_scheduler.ExponentialLR)
    new, bits = gsim.simulate_compression(splats, 25)
    n_s = bits["scales"].shape[0]
    assert 0 < n_s < N // 5 and bits["quats"].shape == (n_s, 4) and bits["sh0"].shape == (n_s, 3)
    new, bits = gsim.simulate_compression(splats, 5)  # before the first entropy step: nothing sampled, no rates
    assert all(b is None for b in bits.values())


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/gsb200.h against _cabi.SIGNATURES, parameter by parameter (count and C type class):
    a drifted ctypes table would otherwise only show up as a wrong argument on the GPU."""
    import ctypes

    from gscodec_studio_b200 import _cabi

    text = open(os.path.join(ROOT, "include", "gsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef struct.*?\}\s*\w+;", "", text, flags=re.S)
    protos = re.findall(r"\b([\w\s\*]+?)\b(gsb200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 40

    def classify(param: str):
        p = " ".join(param.split())
        if p in ("void", ""):
            return None
        if "*" in p or "gsb200_stream_t" in p:
            return ctypes.c_void_p
        base = p.replace("const ", "").split(" ")[0]
        return {"uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "int": ctypes.c_int, "float": ctypes.c_float,
                "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64}[base]

    seen = set()
    for _ret, name, params in protos:
        want = [c for c in (classify(p) for p in params.split(",")) if c is not None]
        got = _cabi.SIGNATURES[name]
        assert len(got) == len(want), (name, len(got), len(want))
        for k, (g, w) in enumerate(zip(got, want)):
            assert g is w, (name, k, g, w)
        seen.add(name)
    assert seen == set(_cabi.SIGNATURES), set(_cabi.SIGNATURES) ^ seen


def test_python_call_sites_pass_the_declared_number_of_arguments():
    """Static check of every `call("gsb200_...", ...)` in the package: the number of positional arguments equals the
    length of the ctypes signature (call sites that splat a tuple are checked for a plausible lower bound)."""
    import ast

    from gscodec_studio_b200 import _cabi

    checked = 0
    roots = [os.path.join(ROOT, "gscodec_studio_b200"), os.path.join(ROOT, "tools")]
    for dirpath, _, files in (item for r in roots for item in os.walk(r)):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "call" and node.args):
                    continue
                first = node.args[0]
                if not (isinstance(first, ast.Constant) and isinstance(first.value, str) and first.value.startswith("gsb200_")):
                    continue
                want = len(_cabi.SIGNATURES[first.value])
                rest = node.args[1:]
                if any(isinstance(a, ast.Starred) for a in rest):
                    assert sum(not isinstance(a, ast.Starred) for a in rest) < want, (f, first.value)
                else:
                    assert len(rest) == want, (f, node.lineno, first.value, len(rest), want)
                checked += 1
    assert checked >= 30


def test_every_kernel_source_is_in_the_build_list():
    from gscodec_studio_b200 import _build

    on_disk = sorted(f for f in os.listdir(_build.CSRC) if f.endswith(".cu"))
    assert on_disk == sorted(_build.SOURCES), (on_disk, _build.SOURCES)


def test_policy_tables_and_checkpoint_keys_match_the_reference_fixture():
    """tests/golden/ref_policy.json was written by oracle/make_golden_policy.py from the REFERENCE's own
    compression_simulation/simulation.py (run on the CPU with a stub _gridencoder): policy tables of both classes, the
    negative-step gate (static only), entropy-model state_dict keys / shapes and optimizer group names."""
    import json

    from gscodec_studio_b200.compression_simulation import simulation as S

    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_policy.json")))
    sim = S.CompressionSimulation(entropy_model_enable=True, entropy_steps=dict(ref["static"]["entropy_steps"]), device="cpu")
    stg = S.STGCompressionSimulation(quantization_sim_type="round", entropy_model_enable=True,
                                     entropy_steps=dict(ref["stg"]["entropy_steps"]), device="cpu")
    for ours, theirs in ((sim, ref["static"]), (stg, ref["stg"])):
        assert ours.simulation_option == theirs["simulation_option"]
        assert ours.q_bitwidth == theirs["q_bitwidth"]
        assert {k: (None if v is None else [float(x) for x in v]) for k, v in ours.bds.items()} == \
               {k: (None if v is None else [float(x) for x in v]) for k, v in theirs["bds"].items()}
        for k, v in theirs["entropy_model_option"].items():  # the reference's dynamic table lists fewer keys
            assert ours.entropy_model_option[k] == v, k
        assert all(not ours.entropy_model_option[k] for k in ours.entropy_model_option if k not in theirs["entropy_model_option"])
        for k, keys in theirs["state_dict"].items():
            got = {n: list(v.shape) for n, v in ours.entropy_models[k].state_dict().items()}
            assert got == keys, k  # same keys, same shapes: either side's checkpoint loads strictly into the other
    assert sim.entropy_min_step == ref["static"]["entropy_min_step"]
    for k, names in ref["static"]["optimizer_param_names"].items():
        assert [g["name"] for g in sim.entropy_model_optimizers[k].param_groups] == names
    # the dynamic policy has no negative-step gate: its rate term is active from step 0 (reference :575-583)
    stg2 = S.STGCompressionSimulation(quantization_sim_type="round", entropy_model_enable=True,
                                      entropy_steps={k: -1 for k in ref["stg"]["entropy_steps"]}, device="cpu")
    assert stg2.entropy_model_option["scales"] and stg2.entropy_model_option["colors"]
    with pytest.raises(KeyError):
        S.CompressionSimulation(entropy_model_enable=True, entropy_steps={"scales": 1}, device="cpu")


def test_gsplat_import_shim():
    """integration/gsplat: `import gsplat` resolves the hot path to this package (the reference's trainers import it
    unchanged); names outside the hot path raise an ImportError that says so (or come from GSPLAT_REFERENCE_PATH)."""
    import subprocess
    import sys

    code = (
        "import gsplat, gscodec_studio_b200 as g\n"
        "assert gsplat.rasterization is g.rasterization and gsplat.isect_tiles is g.isect_tiles\n"
        "from gsplat.rendering import rasterization\n"
        "from gsplat.compression_simulation import CompressionSimulation, STGCompressionSimulation\n"
        "from gsplat.compression_simulation.ops import fake_quantize_ste\n"
        "from gsplat.cuda._wrapper import spherical_harmonics, fully_fused_projection\n"
        "from gsplat.distributed import cli\n"
        "from gsplat.relocation import compute_relocation\n"
        "assert gsplat.rasterization_2dgs is g.rasterization_2dgs\n"
        "from gsplat.cuda._wrapper import fully_fused_projection_2dgs, rasterize_to_pixels_2dgs\n"
        "from gsplat.rendering import rasterization_2dgs\n"
        "from gsplat.utils import depth_to_normal\n"
        "try:\n    gsplat.PngCompression\n    raise SystemExit('the codecs must not resolve')\nexcept ImportError as e:\n    assert 'outside' in str(e)\n"
        "print('shim ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "integration"), ROOT]))
    env.pop("GSPLAT_REFERENCE_PATH", None)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert out.returncode == 0 and "shim ok" in out.stdout, out.stderr[-2000:]


def test_quant_on_load_host_logic_without_kernels(monkeypatch):
    """Host side of the quantise-on-load hook on the CPU (kernels stubbed): simulate_compression(fuse_into_render=True) hands
    back lazy QuantOnLoad objects for the attributes it only fake-quantises, exp / sigmoid of one stay lazy, any other use
    materialises through the stand-alone kernel with the SAME draw, and attributes with a rate term stay eager."""
    from gscodec_studio_b200.compression_simulation import ops, simulation as S

    seen = []

    def fake_apply(x, lo, hi, bits, mode, act, noise, philox):  # stands in for _FakeQuantize.apply
        seen.append((tuple(x.shape), lo, hi, bits, mode, act, noise is not None, philox))
        y = torch.clamp(x, lo, hi) + (noise if noise is not None else 0.0) * (hi - lo) / (2 ** bits - 1)
        return torch.exp(y) if act == 1 else torch.sigmoid(y) if act == 2 else y

    monkeypatch.setattr(ops._FakeQuantize, "apply", staticmethod(fake_apply))
    monkeypatch.setattr(S, "fake_quantize_ste_multi", lambda items: [
        {"output_value": fake_apply(it["input"], it["lower_bd"], it["upper_bd"], it["bitwidth"], 0, 0, it.get("noise"), None),
         "q_step": 0.0} for it in items])
    g = torch.Generator().manual_seed(0)
    N = 64
    splats = {"means": torch.randn(N, 3, generator=g), "scales": torch.randn(N, 3, generator=g), "quats": torch.randn(N, 4, generator=g),
              "opacities": torch.randn(N, generator=g), "sh0": torch.randn(N, 1, 3, generator=g), "shN": torch.randn(N, 15, 3, generator=g)}
    noise = {k: torch.rand(splats[k].shape, generator=g) - 0.5 for k in ("scales", "quats", "opacities", "sh0")}
    sim = S.CompressionSimulation(entropy_model_enable=False, device="cpu")
    # CPU tensors are never made lazy (there is no kernel to consume them): the eager path is taken
    new, bits = sim.simulate_compression(splats, 100, noise=noise, fuse_into_render=True)
    assert all(torch.is_tensor(v) for v in new.values()) and all(b is None for b in bits.values())

    # the lazy object itself (construction needs no kernel when the noise is given)
    lz = ops.QuantOnLoad(splats["scales"], -10, 2, 8, noise["scales"])
    assert lz.shape == (N, 3) and len(lz) == N and lz.dim() == 2 and not seen[4:]
    ex = torch.exp(lz)
    assert isinstance(ex, ops.QuantOnLoad) and ex.act == "exp" and torch.exp(lz) is ex and ex.raw is lz.raw
    d = ex.descriptor()
    assert (d.lo, d.hi, d.bitwidth, d.act) == (-10.0, 2.0, 8, 1) and d.noise == lz.noise.data_ptr()
    n0 = len(seen)
    a = lz + 1.0                       # arithmetic -> materialised once ...
    b = torch.cat([lz, lz], dim=1)     # ... and reused by every later consumer
    assert len(seen) == n0 + 1 and seen[-1][4:7] == (0, 0, True)
    assert torch.equal(a, lz.materialize() + 1.0) and b.shape == (N, 6)
    assert torch.allclose(ex.materialize(), torch.exp(lz.materialize()))
    assert torch.is_tensor(torch.exp(ex)) and ops.solid(lz) is lz.materialize() and ops.solid(None) is None
    with pytest.raises(ValueError):
        ops.QuantOnLoad(splats["scales"], -10, 2, 8, noise["scales"], activation="tanh")


def test_isect_plan_is_host_only_and_sizes_the_private_histogram_rows(monkeypatch):
    """gsb200_isect_plan makes no CUDA call.  With GSB200_ISECT_PRIVATE=1 (CTA-private tile histograms, opt-in) it
    appends one row of C x tiles counters per CTA behind the default layout; the intersection count stays where it was;
    configurations whose histogram does not fit a CTA's shared memory, or that are too small, keep the default size."""
    import ctypes

    from gscodec_studio_b200 import _cabi

    lib = _cabi.load_library()

    def plan(C, n, tw, th):
        n_scratch, total_at = ctypes.c_int64(0), ctypes.c_int64(0)
        assert lib.gsb200_isect_plan(C, n, tw, th, ctypes.byref(n_scratch), ctypes.byref(total_at)) == 0
        return n_scratch.value, total_at.value

    monkeypatch.delenv("GSB200_ISECT", raising=False)
    monkeypatch.setenv("GSB200_ISECT_PRIVATE", "0")
    base, total_at = plan(1, 1_000_000, 120, 68)
    T = 120 * 68
    assert total_at == 33 * T + 1 and base > 33 * T + 2
    monkeypatch.delenv("GSB200_ISECT_PRIVATE")
    assert plan(1, 1_000_000, 120, 68) == (base, total_at)  # opt-in: off by default
    monkeypatch.setenv("GSB200_ISECT_PRIVATE", "1")
    big, total_at1 = plan(1, 1_000_000, 120, 68)
    assert total_at1 == total_at
    rows = big - ((base + 3) & ~3)
    assert rows % T == 0 and rows // T == 2 * 148  # two CTAs per SM
    assert plan(1, 20_000, 120, 68)[0] - ((base + 3) & ~3) == ((20_000 + 511) // 512) * T  # fewer Gaussians, fewer CTAs
    assert plan(1, 1000, 120, 68)[0] == base  # too small: default kernels
    assert plan(8, 1_000_000, 120, 68)[0] == plan(8, 1000, 120, 68)[0]  # 65 280 counters do not fit: default kernels


def test_colour_stream_is_cuda_only_and_opt_out(monkeypatch):
    from gscodec_studio_b200 import rendering

    assert rendering._colour_stream(torch.device("cpu")) is None
    monkeypatch.setenv("GSB200_COLOR_STREAM", "0")
    assert rendering._colour_stream(torch.device("cuda", 0)) is None  # (no CUDA call is made on this path)


def test_blend_kernels_carry_the_blackwell_instructions_the_design_names():
    """SASS census of the shipped stream-blend kernels: the backward stages its record stream with
    the bulk-copy engine (UBLKCP) on mbarriers (SYNCS), both kernels evaluate pixel pairs in packed FP32 (FFMA2 / FMUL2),
    the two-record walk shows as 4 EX2 per trip in both kernels (two records x two pixels), and the headline (D = 3)
    kernels do not spill."""
    import subprocess

    from gscodec_studio_b200 import _cabi

    out = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    funcs = {}
    for block in out.stdout.split("Function : ")[1:]:
        funcs[block.split("\n", 1)[0].strip()] = block
    bwd = [v for k, v in funcs.items() if "blend_stream_bwd_kernelILi3ELi1ELb0" in k]
    fwd = [v for k, v in funcs.items() if "blend_stream_fwd_kernelILi3ELi1E" in k]
    assert len(bwd) == 1 and len(fwd) == 1
    bwd, fwd = bwd[0], fwd[0]
    assert "UBLKCP" in bwd and "SYNCS" in bwd and "SYNCS" in fwd
    for sass in (bwd, fwd):
        assert sass.count("FFMA2") >= 8 and sass.count("FMUL2") >= 8
    assert bwd.count("MUFU.EX2") == 4 and bwd.count("MUFU.RCP") >= 4 and bwd.count("SHFL.BFLY") >= 21
    assert fwd.count("MUFU.EX2") == 4
    assert "STL" not in bwd and "LDL" not in bwd and "STL" not in fwd and "LDL" not in fwd
