"""Node surface: the reference's node classes — same `__call__` schemas — building this package's GPU
objects.  CPU: signatures against the reference's source (when mounted), callback plumbing.  GPU: eight sampler steps
through the nodes' callbacks against fixtures produced by the reference's own node classes (oracle/make_golden.py
--only-nodes), both driven by the scripted sampler of tests/helpers.py."""
import ast
import inspect
import os

import numpy as np
import pytest
import torch

from helpers import Ctx, EngineData, assert_close, scripted_ksampler, t2n

REF = os.environ.get("SRX_REFERENCE_ROOT", "/root/reference")
REF_FILES = {
    "StableRenderSampler": "legacy_codes/nodes/samplers.py",
    "OverlapScheduler": "legacy_codes/nodes/schedulers.py",
    "CorrMapLatentNoiseInitializer": "legacy_codes/nodes/latent.py",
    "DefaultCorresponder": "source/comfyUI/stable_rendering/_nodes/samplers.py",
    "OverlapCorresponder": "source/comfyUI/stable_rendering/_nodes/samplers.py",
    "CorrespondSampler": "source/comfyUI/stable_rendering/_nodes/samplers.py",
}


def _ref_call_schema(path, cls):
    """(parameter names, defaults as source text) of `cls.__call__` in a reference file, without importing it."""
    with open(os.path.join(REF, path)) as f:
        tree = ast.parse(f.read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "__call__":
                    names = [a.arg for a in fn.args.args]
                    defaults = [ast.literal_eval(d) if isinstance(d, ast.Constant) else ast.unparse(d) for d in fn.args.defaults]
                    return names, defaults, getattr(node.body[0], "value", None)
    raise KeyError(cls)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("cls", sorted(REF_FILES))
def test_node_call_schema_matches_reference(cls):
    """Parameter names, their order and their defaults ARE the node schema (types/node_base.py:179-334)."""
    import stable_renderer_b200.nodes as nodes
    names, defaults, _ = _ref_call_schema(REF_FILES[cls], cls)
    sig = inspect.signature(getattr(nodes, cls).__call__)
    ours = list(sig.parameters)
    assert ours == names, f"{cls}: {ours} != {names}"
    ours_defaults = [p.default for p in sig.parameters.values() if p.default is not inspect.Parameter.empty]
    assert len(ours_defaults) == len(defaults)
    for o, r in zip(ours_defaults, defaults):
        if isinstance(r, str) and r in ("_default_sampler", "_default_scheduler"):
            continue                       # the host's first sampler / scheduler name
        assert o == r, f"{cls}: default {o!r} != {r!r}"


def test_legacy_callback_routing_and_timestep_estimate():
    """execute_overlap (samplers.py:79-129): which tensor each option touches, the ddpm-only rule, the timestep estimate."""
    from stable_renderer_b200.nodes import estimated_denoising_timestep, make_overlap_callback
    assert [estimated_denoising_timestep(i, 8) for i in range(8)] == [1000 - int(((i + 1) / 8) * 1000) for i in range(8)]
    calls = []

    def fake_overlap(frame_seq, corr_map, step, timestep):
        calls.append((len(frame_seq), tuple(frame_seq[0].shape), step, timestep))
        return [f + 1 for f in frame_seq]

    for option, sampler, touched in (("noise", "ddpm", (1, 0)), ("denoised", "ddpm", (0, 1)), ("both", "ddpm", (1, 1)),
                                     ("denoised", "ddim", (1, 0)), ("both", "ddim", (1, 0))):
        noise, den = torch.zeros(3, 4, 2, 2), torch.zeros(3, 4, 2, 2)
        ctx = Ctx(noise, step_index=2, total_steps=8, denoised=den)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            make_overlap_callback(fake_overlap, None, option, sampler)(ctx)
        assert (float(noise.mean()), float(den.mean())) == tuple(float(v) for v in touched), (option, sampler)
    assert calls[0] == (3, (1, 4, 2, 2), 2, 1000 - int((3 / 8) * 1000))
    with pytest.raises(ValueError):
        make_overlap_callback(fake_overlap, None, "bogus", "ddpm")(Ctx(torch.zeros(1, 4, 2, 2)))


def test_correspond_sampler_plumbing_on_cpu():
    """CorrespondSampler: ddim/ddpm rule for the overlap corresponder, prepare / step_finished probing, latent fallback."""
    import stable_renderer_b200.nodes as nodes
    from stable_renderer_b200.corresponder import OverlapCorresponder as OC
    nodes.set_ksampler(scripted_ksampler)
    try:
        with pytest.raises(ValueError):
            nodes.CorrespondSampler()(None, None, None, OC(), EngineData(), sampler_name="euler")
        seen = []

        class Probe:
            def prepare(self, engine_data):
                seen.append("prepare")

            def step_finished(self, engine_data, sampling_context):
                seen.append(sampling_context.step_index)
                sampling_context.noise.add_(1.0)

            def finished(self, engine_data, images):
                pass

        ed = EngineData(noise_maps=torch.zeros(2, 4, 2, 2))
        out = nodes.CorrespondSampler()(None, None, None, Probe(), ed, steps=3, sampler_name="ddpm")
        assert seen == ["prepare", 0, 1, 2] and float(out["samples"].mean()) > 0
        corresponder, vae_cb = nodes.OverlapCorresponder()(ed)
        assert isinstance(corresponder, OC) and corresponder.step_finished_inject_ratio == 0.5 and vae_cb(None) is None
        dc, dcb = nodes.DefaultCorresponder()(ed, update_mode="replace")
        assert dc.update_corrmap_mode == "replace" and callable(dcb)
        with pytest.raises(ValueError):
            nodes.CorrespondSampler()(None, None, None, Probe(), None, steps=1, sampler_name="ddpm")
    finally:
        nodes.set_ksampler(None)


@pytest.mark.gpu
@pytest.mark.parametrize("option,sampler", [("noise", "ddpm"), ("denoised", "ddpm"), ("both", "ddpm"), ("denoised", "ddim")])
@pytest.mark.parametrize("algo", ["average", "frame_distance"])
def test_stable_render_sampler_vs_reference_node(golden, option, sampler, algo):
    import warnings
    import stable_renderer_b200.nodes as nodes
    from stable_renderer_b200.overlap import CorrespondenceMap
    g = golden("node_samplers")
    cmap = CorrespondenceMap.from_ids(torch.from_numpy(g["legacy_ids"]).cuda())
    sched = nodes.OverlapScheduler()
    alpha = sched(start_step=1, interpolate_begin=0.9, interpolate_end=0.3, interpolate_type="linear", power=1.0)
    radius = sched(interpolate_begin=0.0, interpolate_end=0.0)
    nodes.set_ksampler(scripted_ksampler)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = nodes.StableRenderSampler()(None, None, None, {"samples": torch.from_numpy(g["legacy_latents"]).cuda()}, cmap,
                                              alpha, radius, overlap_algorithm=algo, apply_overlap_option=option, steps=8,
                                              sampler_name=sampler)
    finally:
        nodes.set_ksampler(None)
    assert_close(t2n(res[0]["samples"]), g[f"legacy_{option}_{sampler}_{algo}"], 5e-5, 1e-5, f"{option}/{sampler}/{algo}")


@pytest.mark.gpu
def test_correspond_sampler_with_overlap_corresponder_vs_reference_node(golden):
    """Eight sampler steps (timesteps 999 ... 124: the last four fall below the injection gate) through the node-built
    corresponder's `step_finished`, in place on the sampler's tensor."""
    import stable_renderer_b200.nodes as nodes
    from stable_renderer_b200.corrmap import IDMap
    g = golden("node_samplers")
    ed = EngineData(id_maps=IDMap(tensor=torch.from_numpy(g["ids"]).cuda()), noise_maps=torch.from_numpy(g["latents"]).cuda())
    nodes.set_ksampler(scripted_ksampler)
    try:
        corresponder, _ = nodes.OverlapCorresponder()(ed, step_finished_inject_ratio=0.5)
        res = nodes.CorrespondSampler()(None, None, None, corresponder, ed, steps=8, sampler_name="ddpm")
    finally:
        nodes.set_ksampler(None)
    assert_close(t2n(res["samples"]), g["current_overlap_ddpm"], 5e-5, 1e-5, "CorrespondSampler + OverlapCorresponder")
