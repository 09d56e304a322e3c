"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the reference's own, unmodified hot-path
files (through oracle/ref_shim.py) on small seeded inputs.  Run in the build container, where /root/reference
is mounted:

    python oracle/make_golden.py

The fixtures pin oracle/srx_oracle.py (tests/test_oracle_vs_golden.py) and serve the `-m gpu` parity tests on
the GPU box, where the reference tree does not exist.  Determinism: torch.set_num_threads(1) — defines the
duplicate-index write order of the reference.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from stable_renderer_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _Ctx:
    def __init__(self, noise, timestep):
        self.noise = noise
        self.denoised = noise
        self.timestep = timestep
        self.step_index = 0
        self.total_steps = 20


class _EngineData:
    def __init__(self, id_maps, correspond_maps=None):
        self.id_maps = id_maps
        self.correspond_maps = correspond_maps


def ref_step(R, ids_t, x_t, ratio, frame_indices=None, timestep=900, stop=500):
    IDMap = R["corrmap"].IDMap
    idm = IDMap(tensor=ids_t.clone(), frame_indices=None if frame_indices is None else list(frame_indices))
    oc = R["corresponder"].OverlapCorresponder(step_finished_inject_ratio=ratio,
                                               step_finished_stop_inject_timestep=stop)
    ctx = _Ctx(x_t.clone(), timestep)
    with ref_shim.quiet():
        oc.step_finished(_EngineData(idm), ctx)
        vsi = idm.create_vertex_screen_info()
    return ctx.noise, vsi, idm.masks


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def step_cases(R):
    cases = {
        # name: (F, H, W, h, w, tex, frac2048, frame_indices, ratio, n_obj)
        "step_sq64_r8": (4, 64, 64, 8, 8, 64, 0.05, None, 0.5, 1),
        "step_sq96_r8_perm": (3, 96, 96, 12, 12, 96, 0.03, [2, 0, 1], 0.1, 1),
        "step_sq100_nonint": (3, 100, 100, 12, 12, 64, 0.0, None, 0.35, 2),
        "step_sq60_to_16": (2, 60, 60, 16, 16, 32, 0.1, None, 1.0, 1),
        "step_dupframe": (3, 64, 64, 8, 8, 64, 0.0, [0, 1, 1], 0.5, 1),
    }
    for name, (F, H, W, h, w, tex, f2048, fidx, ratio, n_obj) in cases.items():
        ids = synthetic.make_ids(F, H, W, tex_h=tex, tex_w=tex, n_obj=n_obj, frac_2048=f2048, seed=11)
        B = F if fidx is None else max(fidx) + 1
        x = synthetic.make_latents(B, 4, h, w, seed=3)
        out, vsi, masks = ref_step(R, ids, x, ratio, fidx)
        save(name, ids=ids.numpy(), x=x.numpy(), ratio=np.float64(ratio),
             frame_indices=np.array(list(range(F)) if fidx is None else fidx, dtype=np.int64),
             out=out.numpy(), vsi=vsi.numpy(), masks=masks.numpy())
    # gate: timestep below stop -> untouched
    ids = synthetic.make_ids(2, 64, 64, tex_h=64, tex_w=64, seed=11)
    x = synthetic.make_latents(2, 4, 8, 8, seed=3)
    out, vsi, _ = ref_step(R, ids, x, 0.5, None, timestep=100, stop=500)
    save("step_gate_off", ids=ids.numpy(), x=x.numpy(), ratio=np.float64(0.5), out=out.numpy(),
         timestep=np.float64(100), stop=np.float64(500))
    # half precision inputs (tolerance cases, corresponder.py:317-318 / math_utils.py:42-51)
    ids = synthetic.make_ids(4, 64, 64, tex_h=64, tex_w=64, seed=11)
    for dt, tag in ((torch.float16, "f16"), (torch.bfloat16, "bf16")):
        x = synthetic.make_latents(4, 4, 8, 8, seed=5, dtype=dt)
        out, _, _ = ref_step(R, ids, x, 0.5)
        save(f"step_half_{tag}", ids=ids.numpy(), x=x.float().numpy(), ratio=np.float64(0.5),
             out=out.float().numpy())
    # crop of the bundled real data (resources/example-sphere-and-object-views/sphere/id, int16, legacy layout)
    d = os.path.join(ref_shim.REFERENCE_ROOT, "resources", "example-sphere-and-object-views", "sphere", "id")
    names = sorted(os.listdir(d), key=lambda n: int(n.split("_")[1].split(".")[0]))[:4]
    crop = np.stack([np.load(os.path.join(d, n))[160:288, 192:320] for n in names])
    ids = torch.from_numpy(crop.copy())
    x = synthetic.make_latents(4, 4, 16, 16, seed=0)
    out, vsi, _ = ref_step(R, ids, x, 0.5)
    save("step_sphere_crop_int16", ids=crop, x=x.numpy(), ratio=np.float64(0.5), out=out.numpy(),
         n_entries=np.int64(vsi.shape[0]))


def bake_cases(R):
    CorrespondMap = R["corrmap"].CorrespondMap
    IDMap = R["corrmap"].IDMap
    F, H, W, tex, k = 4, 64, 64, 32, 2
    ids = synthetic.make_ids(F, H, W, tex_h=tex, tex_w=tex, k=k, frac_2048=0.05, seed=21)
    colors = synthetic.make_colors(F, H, W, 3, seed=9)
    masks = IDMap(tensor=ids.clone()).masks  # 1 = no id
    for mode in ("first", "replace", "first_avg", "replace_avg"):
        cm = CorrespondMap(name="g", k=k, height=tex, width=tex, channel_count=4)
        with ref_shim.quiet():
            cm.update(colors.clone(), ids.clone(), spriteID=1, materialID=0, mode=mode, masks=masks.clone(),
                      inverse_masks=True, ignore_obj_mat_id=True)
        save(f"bake_{mode}_masked", ids=ids.numpy(), colors=colors.numpy(), masks=masks.numpy(),
             values=cm._values.numpy(), writtens=cm._writtens.numpy(), k=np.int64(k), tex=np.int64(tex))
    # two successive updates in 'first' mode (second call must not overwrite), rgba input, no 2048 pixels
    ids2 = synthetic.make_ids(F, H, W, tex_h=tex, tex_w=tex, k=k, n_obj=2, seed=22)
    col4 = synthetic.make_colors(F, H, W, 4, seed=10)
    cm = CorrespondMap(name="g", k=k, height=tex, width=tex, channel_count=4)
    with ref_shim.quiet():
        cm.update(col4[:2].clone(), ids2[:2].clone(), spriteID=2, materialID=0, mode="first")
        cm.update(col4[2:].clone(), ids2[2:].clone(), spriteID=2, materialID=0, mode="first")
    save("bake_first_sprite2_two_calls", ids=ids2.numpy(), colors=col4.numpy(), values=cm._values.numpy(),
         writtens=cm._writtens.numpy(), k=np.int64(k), tex=np.int64(tex), spriteID=np.int64(2))
    # through DefaultCorresponder.finished (corresponder.py:130-155)
    cm = CorrespondMap(name="g", k=k, height=tex, width=tex, channel_count=3)
    dc = R["corresponder"].DefaultCorresponder(update_corrmap_mode="replace", ignore_obj_mat_id_when_update=True)
    ed = _EngineData(IDMap(tensor=ids.clone()), {(1, 0): cm})
    with ref_shim.quiet():
        dc.finished(ed, colors.clone())
    save("bake_finished_replace_c3", ids=ids.numpy(), colors=colors.numpy(), values=cm._values.numpy(),
         writtens=cm._writtens.numpy(), k=np.int64(k), tex=np.int64(tex))


def legacy_cases(R):
    CorrespondenceMap = R["correspondence_map"].CorrespondenceMap
    ov = R["overlap"]
    Scheduler = R["overlap_scheduler"].Scheduler
    factory = R["algorithms"].overlap_algorithm_factory
    T, H, W, h, w = 4, 32, 32, 4, 4
    ids = synthetic.make_ids(T, H, W, tex_h=16, tex_w=16, seed=31, legacy_layout=True, dtype=torch.int16)
    with tempfile.TemporaryDirectory() as td:
        iddir = os.path.join(td, "id")
        os.makedirs(iddir)
        for f in range(T):
            np.save(os.path.join(iddir, f"id_{f}.npy"), ids[f].numpy())
        with ref_shim.quiet():
            cmap = CorrespondenceMap.FromExisting(iddir, enable_cache=False)
        # FromExisting pickles the map beside the id directory (correspondence_map.py:170-172): that file, as written by the
        # reference's own save_cache, is the fixture for the cache loader
        with open(os.path.join(td, "corr_map.pkl"), "rb") as fsrc, open(os.path.join(OUT, "legacy_corr_map.pkl"), "wb") as fdst:
            fdst.write(fsrc.read())
        print(f"legacy_corr_map.pkl: {os.path.getsize(os.path.join(OUT, 'legacy_corr_map.pkl')) / 1024:.1f} KiB")
    keys = np.array(list(cmap.Map.keys()), dtype=np.int64)
    lens = np.array([len(v) for v in cmap.Map.values()], dtype=np.int64)
    flat = np.array([(p[0], p[1], f) for v in cmap.Map.values() for (p, f) in v], dtype=np.int64)
    save("legacy_corrmap", ids=ids.numpy(), keys=keys, lens=lens, traces=flat,
         size=np.array(cmap.size, dtype=np.int64))
    merged = CorrespondenceMap(dict(cmap.Map), cmap.width, cmap.height, cmap.num_frames)
    with ref_shim.quiet():
        merged.merge_nearby(4)
    mkeys = np.array(list(merged.Map.keys()), dtype=np.int64)
    mlens = np.array([len(v) for v in merged.Map.values()], dtype=np.int64)
    mflat = np.array([(p[0], p[1], f) for v in merged.Map.values() for (p, f) in v], dtype=np.int64)
    save("legacy_corrmap_merge4", keys=mkeys, lens=mlens, traces=mflat)

    # dropouts (correspondence_map.py:207-274) on copies of the full and the merged map
    drop = {}
    for tag, base in (("full", cmap), ("merge4", merged)):
        a = CorrespondenceMap(dict(base.Map), base.width, base.height, base.num_frames)
        with ref_shim.quiet():
            a.dropout_index(0.3, 5)
        drop[f"{tag}_index_keys"] = np.array(list(a.Map.keys()), dtype=np.int64)
        b = CorrespondenceMap(dict(base.Map), base.width, base.height, base.num_frames)
        with ref_shim.quiet():
            b.dropout_in_rectangle(((3, 4), (20, 25)), 1)
        drop[f"{tag}_rect_keys"] = np.array(list(b.Map.keys()), dtype=np.int64)
    save("legacy_corrmap_dropout", probability=np.float64(0.3), seed=np.int64(5), rect=np.array([[3, 4], [20, 25]]),
         at_frame=np.int64(1), **drop)

    torch.manual_seed(0)
    frames = [torch.randn(1, 4, h, w, dtype=torch.float64) for _ in range(T)]
    vn = torch.rand(T, H, W, 1, dtype=torch.float64)
    alpha_s = Scheduler(interpolate_begin=0.7, interpolate_end=0.7, interpolate_type="constant")
    res = {}
    for radius in (0, 1):
        rad_s = Scheduler(interpolate_begin=float(radius), interpolate_end=float(radius), interpolate_type="constant")
        for strat in ("average", "frame_distance", "pixel_distance", "perpendicular_view_normal"):
            for cm_name, cm in (("full", cmap), ("merge4", merged)):
                o = ov.ResizeOverlap(alpha_s, rad_s, factory(strat), verbose=False)
                with ref_shim.quiet():
                    outs = o([f.clone() for f in frames], cm, step=0, timestep=500, view_normal_map=vn)
                res[f"out_{strat}_r{radius}_{cm_name}"] = torch.stack(outs).numpy()
    save("legacy_resize_overlap", ids=ids.numpy(), frames=torch.stack(frames).numpy(), view_normal=vn.numpy(),
         alpha=np.float64(0.7), **res)

    # build_view_normal_map (overlap/utils.py:56-102): PIL normal images, [1,3] and [3] view vectors
    from PIL import Image
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 256, size=(3, 8, 10, 3), dtype=np.uint8)
    pil = [Image.fromarray(a, mode="RGB") for a in imgs]
    v13 = torch.tensor([[0.3, -0.5, 0.8]])
    v3 = torch.tensor([0.3, -0.5, 0.8])
    vn13 = R["overlap_utils"].build_view_normal_map(pil, v13).numpy()
    vn3 = R["overlap_utils"].build_view_normal_map(pil, v3).numpy()
    save("legacy_view_normal", images=imgs, v13=v13.numpy(), v3=v3.numpy(), out13=vn13, out3=vn3)

    # scheduler table (overlap_scheduler.py:89-107 / utils.py:24-53); cosine only works for tensor timesteps
    rows = []
    for itype in ("constant", "linear", "exponential", "cosine"):
        for power in (1.0, 2.0):
            s = Scheduler(every_step=2, start_step=2, end_step=40, start_timestep=100, end_timestep=900,
                          interpolate_begin=0.9, interpolate_end=0.2, power=power, interpolate_type=itype,
                          no_interpolate_return=0.05)
            for step in (0, 2, 3, 10, 40, 42):
                for ts in (50, 100, 500, 900, 950):
                    tsv = torch.tensor(float(ts), dtype=torch.float64) if itype == "cosine" else ts
                    v = s(step, tsv)
                    rows.append((("constant", "linear", "exponential", "cosine").index(itype), power, step, ts, float(v)))
    save("legacy_scheduler", table=np.array(rows, dtype=np.float64))


def group_cases(R):
    mu = R["math_utils"]
    t = torch.tensor([[2, 1, 4], [2, 9, 12], [6, 4, 4], [7, 3, 99], [8, 1, 3]])
    a0 = mu.tensor_group_by_then_average(t, index_column=0, value_columns=[1, 2])[0]
    a1, u1 = mu.tensor_group_by_then_average(t, index_column=1, value_columns=[0], return_unique=True)
    g = torch.Generator().manual_seed(1)
    big = torch.cat([torch.randn(5000, 4, generator=g), torch.randint(0, 37, (5000, 1), generator=g).float()], dim=1)
    b, ub = mu.tensor_group_by_then_average(big, index_column=-1, value_columns=[0, 1, 2, 3], return_unique=True)
    c = torch.randn(3, 4, 8, 8, generator=g)
    s = torch.randn(3, 4, 8, 8, generator=g) * 2 + 1
    ad = mu.adaptive_instance_normalization(c, s)
    save("group_by_average", t=t.numpy(), a0=a0.numpy(), a1=a1.numpy(), u1=u1.numpy(), big=big.numpy(),
         b=b.numpy(), ub=ub.numpy(), content=c.numpy(), style=s.numpy(), adain=ad.numpy())


def randn_init_cases(R):
    """tensor_group_by_then_randn_init (math_utils.py:164-229) on the CPU generator, and the arithmetic of
    CreateNoiseSequenceFromIdMap (_nodes/loaders.py:193-271).  The node class itself cannot be imported (it needs the
    ComfyUI type system), so its body is replayed here step by step around the reference's REAL grouping function, on CPU."""
    import torch.nn.functional as Fnn
    mu = R["math_utils"]
    g = torch.Generator().manual_seed(3)
    t = torch.cat([torch.randn(4000, 3, generator=g), torch.randint(0, 500, (4000, 1), generator=g).float()], dim=1)
    torch.manual_seed(11)
    exp, uniq = mu.tensor_group_by_then_randn_init(t, index_column=-1, value_columns=[0, 1, 2], return_unique=True)
    torch.manual_seed(11)
    table = torch.randn(len(uniq), 3)
    u2, inv = t[:, -1].unique(return_inverse=True)
    assert torch.equal(u2, uniq) and torch.equal(exp, table[inv]), "randn_like(expanded unique) != randn(n_unique, C)"
    save("randn_init", t=t.numpy(), expanded=exp.numpy(), unique=uniq.numpy(), inverse=inv.numpy().astype(np.int32),
         table=table.numpy(), seed=np.int64(11))

    # CreateNoiseSequenceFromIdMap: the reference's own __call__ body (ref_shim.noise_node_call), on CPU
    import hashlib
    from stable_renderer_b200 import synthetic
    corrmap = R["corrmap"]
    node_call = ref_shim.noise_node_call()
    size, F, seed = 512, 1, 77
    ids = synthetic.make_ids(F, size, size, tex_h=96, tex_w=96, frac_2048=0.05, seed=17)
    out = {}
    for option in ("nearest", "mean", "max", "min"):
        with ref_shim.quiet():
            res = node_call(None, corrmap.IDMap(tensor=ids.clone()), seed, "SD15", option)
        if option == "nearest":
            out["nearest_samples"], out["nearest_noise"] = res["samples"].numpy(), res["noise"].numpy()
        else:
            assert not res["samples"].any()
            out[option + "_noise"] = res["noise"].numpy()
    save("noise_from_idmap", seed=np.int64(seed), id_seed=np.int64(17), tex=np.int64(96), size=np.int64(size), frames=np.int64(F), **out)

    # id maps of another size than the node's working size, permuted / shared latent frames: digest + every 61st value of the
    # reference node's outputs (the full tensors are megabytes of random floats)
    more = {}
    names = []
    for case, (H, fi) in {"map_1024_on_sd15": (1024, [0, 1]), "map_256_on_sd15": (256, [1, 0]), "map_384_on_sd15": (384, [0, 1]),
                          "two_id_frames_one_latent_frame": (512, [0, 0, 2]), "permuted_frames": (512, [2, 0, 1])}.items():
        ids_c = synthetic.make_ids(len(fi), H, H, tex_h=150, tex_w=150, frac_2048=0.1, seed=9)
        names.append(f"{case}:{H}:{','.join(map(str, fi))}")
        for option in ("nearest", "max", "mean"):
            with ref_shim.quiet():
                res = node_call(None, corrmap.IDMap(tensor=ids_c.clone(), frame_indices=list(fi)), 321, "SD15", option)
            for k in (("samples", "noise") if option == "nearest" else ("noise",)):
                arr = np.ascontiguousarray(res[k].numpy())
                more[f"{case}_{option}_{k}_shape"] = np.array(arr.shape, dtype=np.int64)
                more[f"{case}_{option}_{k}_every61"] = arr.reshape(-1)[::61].copy()
                more[f"{case}_{option}_{k}_sha256"] = np.array(hashlib.sha256(arr.tobytes()).hexdigest())
    save("noise_node_other_sizes", cases=np.array(names), seed=np.int64(321), id_seed=np.int64(9), tex=np.int64(150), **more)


def dump_cases(R):
    """CorrespondMap.dump / Load (corrmap.py:738-872) on a small atlas with values outside [0,1] and halves that round."""
    import tempfile
    from PIL import Image
    cmod = R["corrmap"]
    m = cmod.CorrespondMap(name="t", k=2, height=8, width=16, channel_count=4)
    g = torch.Generator().manual_seed(0)
    m._values[:] = (torch.rand(4, 128, 4, generator=g) * 1.4 - 0.2).half()
    m._writtens[:] = torch.rand(4, 128, generator=g) > 0.5
    d = tempfile.mkdtemp()
    with ref_shim.quiet():
        p = m.dump(d, name="abc")
        m2 = cmod.CorrespondMap.Load(p)
    imgs = np.stack([np.array(Image.open(os.path.join(p, f"{i}.png"))) for i in range(4)])
    flags = np.stack([np.array(Image.open(os.path.join(p, f"{i}_written.png"))) for i in range(4)])
    meta = open(os.path.join(p, "meta.json")).read()
    save("corrmap_dump", values=m._values.numpy().view(np.uint16), writtens=m._writtens.numpy(), png=imgs, png_written=flags,
         loaded_values=m2._values.numpy().view(np.uint16), loaded_writtens=m2._writtens.numpy(), meta=np.array(meta))


def latent_init_cases(R):
    """The legacy node CorrMapLatentNoiseInitializer.__call__ (legacy_codes/nodes/latent.py:10-40), the real class, on the
    `legacy_corrmap` id buffers: full map and merge_nearby(4); square 8x down-sample, a non-integer ratio, batch > frames."""
    CorrespondenceMap = R["correspondence_map"].CorrespondenceMap
    Node = R["legacy_latent_node"].CorrMapLatentNoiseInitializer
    T, H, W = 4, 32, 32
    ids = synthetic.make_ids(T, H, W, tex_h=16, tex_w=16, seed=31, legacy_layout=True, dtype=torch.int16)
    with tempfile.TemporaryDirectory() as td:
        iddir = os.path.join(td, "id")
        os.makedirs(iddir)
        for f in range(T):
            np.save(os.path.join(iddir, f"id_{f}.npy"), ids[f].numpy())
        with ref_shim.quiet():
            cmap = CorrespondenceMap.FromExisting(iddir, enable_cache=False)
    merged = CorrespondenceMap(dict(cmap.Map), cmap.width, cmap.height, cmap.num_frames)
    with ref_shim.quiet():
        merged.merge_nearby(4)
    out = {}
    cases = []
    for tag, cm in (("full", cmap), ("merge4", merged)):
        for (width, height, batch, seed) in ((256, 256, 4, 7), (80, 48, 6, 123456789), (64, 64, 4, 0)):
            with ref_shim.quiet():
                (d,) = Node()(width, height, batch, seed, cm)
            name = f"{tag}_{width}x{height}_b{batch}_s{seed}"
            out[name + "_samples"] = d["samples"].numpy()
            out[name + "_noise"] = d["noise"].numpy()
            cases.append((tag, width, height, batch, seed))
        out[tag + "_n_traces"] = np.int64(sum(1 for v in cm.Map.values() if len(v) > 1))
    save("legacy_latent_init", ids=ids.numpy(), cases=np.array([f"{t}:{w}:{h}:{b}:{s}" for t, w, h, b, s in cases]), **out)


def _gbuffer_attachments(g, H, W, coverage):
    """Synthetic attachments in GL row order (origin bottom-left), dtypes of renderManager.py:206-367."""
    alpha = (torch.rand(H, W, generator=g) < coverage).float()
    edge = torch.rand(H, W, generator=g) < 0.1
    alpha = torch.where(edge, torch.rand(H, W, generator=g), alpha)              # anti-aliased edges: fractional coverage
    color = torch.cat([torch.rand(H, W, 3, generator=g), alpha.unsqueeze(-1)], dim=-1).half()
    ids = torch.randint(0, 5000, (H, W, 4), generator=g, dtype=torch.int32) * (alpha > 0).int().unsqueeze(-1)
    pos = torch.randn(H, W, 3, generator=g)
    nd = torch.cat([torch.rand(H, W, 3, generator=g), (torch.rand(H, W, 1, generator=g) * (alpha > 0).float().unsqueeze(-1))], dim=-1).half()
    noise = (torch.randn(H, W, 4, generator=g) * 1.3 + 0.2).half()
    canny = torch.rand(H, W, 3, generator=g).half()        # cannyFBOTex is read as a HALF tensor (renderManager.py:353)
    return dict(color=color, ids=ids, pos=pos, normal_depth=nd, noise=noise, canny=canny)


class _FboTex:
    """Stand-in for a G-buffer `Texture`: holds the attachment in GL row order; `tensor(update, flip)` is texture.py:221-254 for a
    CPU tensor (the flip is the only arithmetic of that method)."""

    def __init__(self, data=None):
        self.data = data

    def tensor(self, update=True, flip=True):
        return self.data.flip(0) if flip else self.data


class _Ns:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def ingest_cases(R):
    """RenderManager._save_frame_data (renderManager.py:877-948) and the closer-pixel merge inside _wrapIdenticalGBufferTask
    (:88-133): the reference's OWN function bodies (ref_shim.render_manager_functions cuts them out of the file; the module cannot
    be imported — OpenGL, window, managers) run against stand-ins for the manager and its seven textures."""
    save_frame_data, wrap_task = ref_shim.render_manager_functions()
    g = torch.Generator().manual_seed(2024)
    H, W = 48, 64
    bg = torch.randn((1, H, W, 4), generator=g, dtype=torch.float32)
    tex_names = dict(color="colorFBOTex", ids="idFBOTex", pos="posFBOTex", normal_depth="normal_and_depth_FBOTex",
                     noise="noiseFBOTex", canny="cannyFBOTex")
    rm = _Ns(data_to_be_added_to_engineData={}, GlobalBGNoise=bg, engine=_Ns(RuntimeManager=_Ns(FrameCount=0)),
             **{v: _FboTex() for v in tex_names.values()})
    src_all = []
    for frame in range(2):
        src = _gbuffer_attachments(g, H, W, 0.55)
        src_all.append(src)
        for k, v in src.items():
            getattr(rm, tex_names[k]).data = v
        rm.engine.RuntimeManager.FrameCount = frame
        save_frame_data(rm)
    data = rm.data_to_be_added_to_engineData
    assert data.pop("frame_indices") == [0, 1]
    out = {"bg_noise": bg.numpy()}
    for f, src in enumerate(src_all):
        for k, v in src.items():
            out[f"src{f}_{k}"] = v.numpy().view(np.uint16) if v.dtype == torch.float16 else v.numpy()
    for k, v in data.items():
        out[k] = v.numpy().view(np.uint16) if v.dtype == torch.float16 else v.numpy()

    # closer-pixel merge of three draws with the temp buffers of renderManager.py:219-357
    rm._color_buffer_temp = torch.zeros(H, W, 4, dtype=torch.float16)
    rm._id_buffer_temp = torch.zeros(H, W, 4, dtype=torch.int32)
    rm._pos_buffer_temp = torch.zeros(H, W, 3)
    rm._normal_buffer_temp = torch.zeros(H, W, 3, dtype=torch.float16)
    rm._depth_buffer_temp = torch.zeros(H, W, dtype=torch.float16)
    rm._noise_buffer_temp = torch.zeros(H, W, 4, dtype=torch.float16)
    rm._canny_buffer_temp = torch.zeros(H, W, 3, dtype=torch.float16)
    rm.BindGBufferTexToShader = lambda shader: None
    rm._update_gbuffer_tex_to_shader_binding_tex = lambda: None
    shader = _Ns(useProgram=lambda: None)
    for d in range(3):
        src = _gbuffer_attachments(g, H, W, 0.4)
        if d == 2:
            src["normal_depth"][..., 3] = src_prev["normal_depth"][..., 3]                  # equal depths: the earlier draw stays
        src_prev = src
        for k, v in src.items():
            out[f"draw{d}_{k}"] = v.numpy().view(np.uint16) if v.dtype == torch.float16 else v.numpy()

        def draw(src=src):                                                                    # the "task": the draw fills the attachments
            for k, v in src.items():
                getattr(rm, tex_names[k]).data = v
        wrap_task(rm, draw, shader, None, None, True)
    temp = dict(color=rm._color_buffer_temp, ids=rm._id_buffer_temp, pos=rm._pos_buffer_temp, normal=rm._normal_buffer_temp,
                depth=rm._depth_buffer_temp, noise=rm._noise_buffer_temp, canny=rm._canny_buffer_temp)
    for k, v in temp.items():
        out[f"temp_{k}"] = v.numpy().view(np.uint16) if v.dtype == torch.float16 else v.numpy()
    out["generator"] = np.array("RenderManager._save_frame_data and _wrapIdenticalGBufferTask bodies of the reference, unmodified, on stand-in textures")
    save("frame_ingest", **out)


def node_cases():
    """The reference's node classes (legacy StableRenderSampler / OverlapScheduler, current CorrespondSampler + OverlapCorresponder
    node) driven by the scripted sampler of tests/helpers.py: 8 sampler steps through the nodes' own callbacks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import scripted_ksampler
    N = ref_shim.load_reference_nodes(scripted_ksampler)
    R = ref_shim.load_reference()
    # -- legacy: StableRenderSampler.execute_overlap (legacy_codes/nodes/samplers.py:79-129)
    CorrespondenceMap = R["correspondence_map"].CorrespondenceMap
    T, H, W, h, w = 4, 32, 32, 4, 4
    ids = synthetic.make_ids(T, H, W, tex_h=16, tex_w=16, seed=41, legacy_layout=True, dtype=torch.int16)
    with tempfile.TemporaryDirectory() as td:
        iddir = os.path.join(td, "id")
        os.makedirs(iddir)
        for f in range(T):
            np.save(os.path.join(iddir, f"id_{f}.npy"), ids[f].numpy())
        with ref_shim.quiet():
            cmap = CorrespondenceMap.FromExisting(iddir, enable_cache=False)
    torch.manual_seed(3)
    lat = torch.randn(T, 4, h, w)
    sched = N["legacy_schedulers"].OverlapScheduler()
    alpha = sched(start_step=1, interpolate_begin=0.9, interpolate_end=0.3, interpolate_type="linear", power=1.0)
    radius = sched(interpolate_begin=0.0, interpolate_end=0.0)
    node = N["legacy_samplers"].StableRenderSampler()
    out = {}
    for option, sampler in (("noise", "ddpm"), ("denoised", "ddpm"), ("both", "ddpm"), ("denoised", "ddim")):
        for algo in ("average", "frame_distance"):
            with ref_shim.quiet():
                res = node(None, None, None, {"samples": lat.clone()}, cmap, alpha, radius, overlap_algorithm=algo,
                           apply_overlap_option=option, steps=8, sampler_name=sampler)
            out[f"legacy_{option}_{sampler}_{algo}"] = res[0]["samples"].numpy()
    # -- current generation: CorrespondSampler + the OverlapCorresponder node (_nodes/samplers.py:71-201)
    F, Hc, hc = 4, 64, 8
    ids_c = synthetic.make_ids(F, Hc, Hc, tex_h=32, tex_w=32, frac_2048=0.05, seed=43)
    x = torch.randn(F, 4, hc, hc)
    ed = _EngineData(R["corrmap"].IDMap(tensor=ids_c.clone()))
    ed.noise_maps = x.clone()
    with ref_shim.quiet():
        corresponder, vae_cb = N["samplers"].OverlapCorresponder()(ed, step_finished_inject_ratio=0.5)
        res = N["samplers"].CorrespondSampler()(None, None, None, corresponder, ed, steps=8, sampler_name="ddpm")
    out["current_overlap_ddpm"] = res["samples"].numpy()
    save("node_samplers", legacy_ids=ids.numpy(), legacy_latents=lat.numpy(), ids=ids_c.numpy(), latents=x.numpy(), **out)


def feature_cases(R):
    """§8f-4: the post_atten_inject body (early return bypassed, ref_shim.post_atten_inject_body) on [B, hw, c] features with
    the reference's own (quirky) up-sampling size (IDMap.height, IDMap.width) = (W, 4), and taichi_cells_overlap run as Python."""
    body = ref_shim.post_atten_inject_body()
    IDMap = R["corrmap"].IDMap
    out = {}
    g = torch.Generator().manual_seed(11)
    for tag, (F, H, hw, c) in (("a", (3, 64, 64, 40)), ("b", (2, 96, 144, 24))):
        ids = synthetic.make_ids(F, H, H, tex_h=32, tex_w=32, frac_2048=0.05, seed=51)
        feats = torch.randn(F, hw, c, generator=g)
        oc = R["corresponder"].OverlapCorresponder(post_attn_inject_ratio=0.6)
        with ref_shim.quiet():
            res = body(oc, None, _EngineData(IDMap(tensor=ids.clone())), feats.clone(), 12)
        out[f"ids_{tag}"], out[f"features_{tag}"], out[f"out_{tag}"] = ids.numpy(), feats.numpy(), res.numpy()
    # cell-similarity overlap: 2 frames of 8x8 pixels, 16 cells of 4 consecutive pixels, 6-wide values
    fn = ref_shim.taichi_cells_overlap_python()
    idc = synthetic.make_ids(2, 8, 8, tex_h=3, tex_w=3, seed=52).reshape(2, 64, 4).float()
    vals = torch.randn(2, 16, 6, generator=g)
    contrib = torch.rand(2, 64, generator=g) * 0.25
    new = torch.zeros_like(vals)
    fn(idc, vals, new, contrib)
    out["cells_ids"], out["cells_values"], out["cells_contrib"], out["cells_new"] = idc.numpy(), vals.numpy(), contrib.numpy(), new.numpy()
    save("feature_overlap", ratio=np.float64(0.6), **out)


def johnny_cases(R):
    """L7: johnny_overlap.overlap (ref_shim.johnny_overlap_function: the one broken statement replaced) on a 4-frame map."""
    fn = ref_shim.johnny_overlap_function()
    CorrespondenceMap = R["correspondence_map"].CorrespondenceMap
    T, H, W, h, w = 4, 32, 32, 4, 4
    ids = synthetic.make_ids(T, H, W, tex_h=16, tex_w=16, seed=71, legacy_layout=True, dtype=torch.int16)
    with tempfile.TemporaryDirectory() as td:
        iddir = os.path.join(td, "id")
        os.makedirs(iddir)
        for f in range(T):
            np.save(os.path.join(iddir, f"id_{f}.npy"), ids[f].numpy())
        with ref_shim.quiet():
            cmap = CorrespondenceMap.FromExisting(iddir, enable_cache=False)
    assert len(cmap.Map) >= 100, "the reference's progress bar divides by len(Map) // 100"
    g = torch.Generator().manual_seed(9)
    frames = [torch.randn(1, 4, h, w, generator=g, dtype=torch.float64) for _ in range(T)]
    orig = [torch.randn(1, 4, h, w, generator=g, dtype=torch.float64) for _ in range(T)]
    noise = [torch.randn(1, 4, h, w, generator=g, dtype=torch.float64) for _ in range(T)]

    class _Sched:
        @staticmethod
        def add_noise(lat, nz, t):
            return lat * 0.8 + nz * 0.6

    class _Pipe:
        scheduler = _Sched()

    out = {}
    with ref_shim.quiet():
        out["out_beta0"] = torch.stack(fn([f.clone() for f in frames], cmap, _Pipe(), step=3, timestep=500)).numpy()
        out["out_beta03"] = torch.stack(fn([f.clone() for f in frames], cmap, _Pipe(), step=3, timestep=500, init_latents_orig_seq=orig,
                                           noise_seq=noise, beta=0.3)).numpy()
        gated = fn([f.clone() for f in frames], cmap, _Pipe(), step=3, timestep=1500)          # timestep > start_timestep: alpha = 0
    out["gated_is_input"] = np.array(all(torch.equal(a, b) for a, b in zip(gated, frames)))
    save("johnny_overlap", ids=ids.numpy(), frames=torch.stack(frames).numpy(), orig=torch.stack(orig).numpy(),
         noise=torch.stack(noise).numpy(), **out)


def interp_cases(R):
    """ResizeOverlap with interpolate_mode != 'nearest' (overlap.py:205-221): the latents really are resampled."""
    CorrespondenceMap = R["correspondence_map"].CorrespondenceMap
    ov = R["overlap"]
    Scheduler = R["overlap_scheduler"].Scheduler
    factory = R["algorithms"].overlap_algorithm_factory
    T, H, W, h, w = 4, 32, 32, 4, 4
    ids = synthetic.make_ids(T, H, W, tex_h=16, tex_w=16, seed=31, legacy_layout=True, dtype=torch.int16)
    with tempfile.TemporaryDirectory() as td:
        iddir = os.path.join(td, "id")
        os.makedirs(iddir)
        for f in range(T):
            np.save(os.path.join(iddir, f"id_{f}.npy"), ids[f].numpy())
        with ref_shim.quiet():
            cmap = CorrespondenceMap.FromExisting(iddir, enable_cache=False)
    torch.manual_seed(5)
    frames = [torch.randn(1, 4, h, w, dtype=torch.float32) for _ in range(T)]
    alpha_s = Scheduler(interpolate_begin=0.7, interpolate_end=0.7, interpolate_type="constant")
    rad_s = Scheduler(interpolate_begin=0.0, interpolate_end=0.0, interpolate_type="constant")
    res = {}
    for mode in ("bilinear", "bicubic", "area"):
        for strat in ("average", "frame_distance"):
            o = ov.ResizeOverlap(alpha_s, rad_s, factory(strat), verbose=False, interpolate_mode=mode)
            with ref_shim.quiet():
                outs = o([f.clone() for f in frames], cmap, step=0, timestep=500)
            res[f"out_{mode}_{strat}"] = torch.stack(outs).numpy()
    save("legacy_resize_overlap_interp", ids=ids.numpy(), frames=torch.stack(frames).numpy(), alpha=np.float64(0.7), **res)


def main():
    if "--only-ingest" in sys.argv:
        torch.set_num_threads(1)
        ingest_cases(ref_shim.load_reference())
        return
    if "--only-randn" in sys.argv:
        torch.set_num_threads(1)
        randn_init_cases(ref_shim.load_reference())
        return
    if "--only-legacy" in sys.argv:
        torch.set_num_threads(1)
        legacy_cases(ref_shim.load_reference())
        return
    if "--only-latent-init" in sys.argv:
        torch.set_num_threads(1)
        latent_init_cases(ref_shim.load_reference())
        return
    if "--only-interp" in sys.argv:
        torch.set_num_threads(1)
        interp_cases(ref_shim.load_reference())
        return
    if "--only-johnny" in sys.argv:
        torch.set_num_threads(1)
        johnny_cases(ref_shim.load_reference())
        return
    if "--only-features" in sys.argv:
        torch.set_num_threads(1)
        feature_cases(ref_shim.load_reference())
        return
    if "--only-nodes" in sys.argv:
        torch.set_num_threads(1)
        node_cases()
        return
    if not ref_shim.available():
        raise SystemExit("reference tree not mounted; fixtures can only be regenerated in the build container")
    torch.set_num_threads(1)
    R = ref_shim.load_reference()
    group_cases(R)
    randn_init_cases(R)
    dump_cases(R)
    step_cases(R)
    bake_cases(R)
    legacy_cases(R)
    latent_init_cases(R)
    ingest_cases(R)
    feature_cases(R)
    johnny_cases(R)
    interp_cases(R)
    node_cases()


if __name__ == "__main__":
    main()
