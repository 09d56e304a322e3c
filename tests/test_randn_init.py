"""Same-key broadcast initialisers: `tensor_group_by_then_randn_init` and
`CreateNoiseSequenceFromIdMap`.  CPU tests pin the numpy oracle to the reference-generated fixtures; -m gpu tests compare
the CUDA path with the fixtures (per-key rows drawn from the seeded CPU generator, as in the fixture run) and the oracle."""
import numpy as np
import pytest
import torch

import srx_oracle as O


def _node_draws(seed, size, n_unique):
    """The node's RNG calls in its order (loaders.py:207-219 + the two draws inside tensor_group_by_then_randn_init), on CPU."""
    g1 = torch.manual_seed(seed)
    g2 = torch.manual_seed(seed + 1)
    base_latent = torch.randn([1, 4, size, size], device="cpu", generator=g1)
    base_noise = torch.randn([1, 4, size, size], device="cpu", generator=g2)
    key_latent = torch.randn(n_unique, 4)
    key_noise = torch.randn(n_unique, 4)
    return [t.numpy() for t in (base_latent, base_noise, key_latent, key_noise)]


def _node_ids(g):
    from stable_renderer_b200 import synthetic
    return synthetic.make_ids(int(g["frames"]), int(g["size"]), int(g["size"]), tex_h=int(g["tex"]), tex_w=int(g["tex"]),
                              frac_2048=0.05, seed=int(g["id_seed"]))


def test_oracle_randn_init_matches_reference(golden):
    g = golden("randn_init")
    exp, uniq = O.randn_init_expand(g["t"][:, -1], g["table"])
    assert np.array_equal(exp, g["expanded"]) and np.array_equal(uniq, g["unique"])
    assert np.array_equal(O.group_slots(g["t"][:, -1])[1], g["inverse"])


def test_oracle_noise_sequence_matches_reference(golden):
    g = golden("noise_from_idmap")
    ids = _node_ids(g).numpy()
    size = int(g["size"])
    n_unique = len(np.unique(O.vertex_screen_info(ids, None)[:, 3]))
    bl, bn, kl, kn = _node_draws(int(g["seed"]), size, n_unique)
    s, n = O.noise_sequence_from_ids(ids, None, bl, bn, kl, kn, size, "nearest")
    assert np.array_equal(s, g["nearest_samples"]) and np.array_equal(n, g["nearest_noise"])
    for option in ("max", "min"):
        s, n = O.noise_sequence_from_ids(ids, None, bl, bn, kl, kn, size, option)
        assert np.array_equal(n, g[option + "_noise"]) and not s.any()
    _, n = O.noise_sequence_from_ids(ids, None, bl, bn, kl, kn, size, "mean")
    assert n.shape == g["mean_noise"].shape == (2, 4, 64, 64)        # the node's view arithmetic doubles the frames
    np.testing.assert_allclose(n, g["mean_noise"], rtol=1e-6, atol=1e-7)


def _other_cases(g):
    for entry in g["cases"]:
        case, H, fi = str(entry).split(":")
        yield case, int(H), [int(v) for v in fi.split(",")]


def _check_against_node_fixture(g, case, option, key, arr):
    """Fixture = digest + every 61st value of the reference node's output (oracle/make_golden.py::randn_init_cases)."""
    import hashlib
    arr = np.ascontiguousarray(arr)
    assert tuple(g[f"{case}_{option}_{key}_shape"]) == arr.shape, (case, option, key)
    if option == "mean":                                  # float sums in another order
        np.testing.assert_allclose(arr.reshape(-1)[::61], g[f"{case}_{option}_{key}_every61"], rtol=1e-6, atol=1e-7)
    else:
        assert np.array_equal(arr.reshape(-1)[::61], g[f"{case}_{option}_{key}_every61"]), (case, option, key)
        assert hashlib.sha256(arr.tobytes()).hexdigest() == str(g[f"{case}_{option}_{key}_sha256"]), (case, option, key)


def test_oracle_noise_sequence_other_sizes_matches_reference_node(golden):
    """Id maps of another size than the working size (1024 / 256 / 384 on SD15), two id frames on one latent frame, permuted
    frames: the oracle against the reference node's own __call__ body, bit for bit (digests of the full outputs)."""
    from stable_renderer_b200 import synthetic
    g = golden("noise_node_other_sizes")
    size, seed = 512, int(g["seed"])
    for case, H, fi in _other_cases(g):
        ids = synthetic.make_ids(len(fi), H, H, tex_h=int(g["tex"]), tex_w=int(g["tex"]), frac_2048=0.1, seed=int(g["id_seed"])).numpy()
        n_unique = len(np.unique(O.vertex_screen_info(ids, fi)[:, 3]))
        bl, bn, kl, kn = _node_draws(seed, size, n_unique)
        for option in ("nearest", "max", "mean"):
            s_, n_ = O.noise_sequence_from_ids(ids, fi, bl, bn, kl, kn, size, option)
            _check_against_node_fixture(g, case, option, "noise", n_)
            if option == "nearest":
                _check_against_node_fixture(g, case, option, "samples", s_)


@pytest.mark.gpu
def test_gpu_noise_sequence_other_sizes_matches_reference_node(golden):
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.loaders import CreateNoiseSequenceFromIdMap
    g = golden("noise_node_other_sizes")
    for case, H, fi in _other_cases(g):
        ids = synthetic.make_ids(len(fi), H, H, tex_h=int(g["tex"]), tex_w=int(g["tex"]), frac_2048=0.1, seed=int(g["id_seed"]))
        idm = IDMap(tensor=ids.cuda(), frame_indices=fi)
        for option in ("nearest", "max", "mean"):
            out = CreateNoiseSequenceFromIdMap()(idm, int(g["seed"]), "SD15", option, rng_device="cpu")
            _check_against_node_fixture(g, case, option, "noise", out["noise"].cpu().numpy())
            if option == "nearest":
                _check_against_node_fixture(g, case, option, "samples", out["samples"].cpu().numpy())


@pytest.mark.gpu
def test_gpu_randn_init_matches_reference_grouping(golden):
    from stable_renderer_b200.math_utils import tensor_group_by_then_randn_init
    g = golden("randn_init")
    t = torch.from_numpy(g["t"]).cuda()
    torch.manual_seed(5)
    out, uniq = tensor_group_by_then_randn_init(t, index_column=-1, value_columns=[0, 1, 2], return_unique=True)
    torch.manual_seed(5)
    table = torch.randn(len(g["unique"]), 3, device="cuda")          # the reference's draw on this device
    inv = torch.from_numpy(g["inverse"]).long().cuda()
    assert torch.equal(uniq.cpu(), torch.from_numpy(g["unique"]))    # sorted unique keys: bit exact
    assert torch.equal(out, table[inv])                              # expansion through the reference's inverse: bit exact
    (only,) = tensor_group_by_then_randn_init(t, index_column=-1, value_columns=[0])
    assert only.shape == (t.shape[0], 1)
    with pytest.raises(ValueError):
        tensor_group_by_then_randn_init(t, index_column=9, value_columns=[0])
    from stable_renderer_b200 import _lib
    with pytest.raises(_lib.SrxUnavailable):
        tensor_group_by_then_randn_init(t.cpu(), index_column=-1, value_columns=[0])


@pytest.mark.gpu
def test_gpu_noise_sequence_matches_reference(golden):
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.loaders import CreateNoiseSequenceFromIdMap
    g = golden("noise_from_idmap")
    ids = _node_ids(g)
    node = CreateNoiseSequenceFromIdMap()
    for id_dtype in (torch.int32, torch.int16):
        idm = IDMap(tensor=ids.to(id_dtype).cuda())
        out = node(idm, int(g["seed"]), "SD15", "nearest", rng_device="cpu")
        assert np.array_equal(out["samples"].cpu().numpy(), g["nearest_samples"])
        assert np.array_equal(out["noise"].cpu().numpy(), g["nearest_noise"])
        for option in ("max", "min"):
            out = node(idm, int(g["seed"]), "SD15", option, rng_device="cpu")
            assert np.array_equal(out["noise"].cpu().numpy(), g[option + "_noise"])
            assert not out["samples"].any()
        out = node(idm, int(g["seed"]), "SD15", "mean", rng_device="cpu")
        np.testing.assert_allclose(out["noise"].cpu().numpy(), g["mean_noise"], rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        node(idm, 1, "SD3")
    with pytest.raises(ValueError):
        node(idm, 1, "SD15", "median")


@pytest.mark.gpu
def test_gpu_noise_sequence_permuted_frames_vs_oracle():
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.loaders import CreateNoiseSequenceFromIdMap
    size, F, seed = 512, 3, 123
    ids = synthetic.make_ids(F, size, size, tex_h=200, tex_w=200, frac_2048=0.1, seed=4)
    fi = [2, 0, 1]
    n_unique = len(np.unique(O.vertex_screen_info(ids.numpy(), fi)[:, 3]))
    bl, bn, kl, kn = _node_draws(seed, size, n_unique)
    want_s, want_n = O.noise_sequence_from_ids(ids.numpy(), fi, bl, bn, kl, kn, size, "nearest")
    out = CreateNoiseSequenceFromIdMap()(IDMap(tensor=ids.cuda(), frame_indices=fi), seed, "SD15", "nearest", rng_device="cpu")
    assert np.array_equal(out["samples"].cpu().numpy(), want_s) and np.array_equal(out["noise"].cpu().numpy(), want_n)
    # same texel => same start vector in every frame that sees it (the point of the node)
    vsi = O.vertex_screen_info(ids.numpy(), fi)
    k0 = vsi[(vsi[:, 4] * size).astype(int) % 8 == 0]
    assert k0.size > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["map_1024_on_sd15", "map_256_on_sd15", "map_384_on_sd15", "two_id_frames_one_latent_frame"])
def test_gpu_noise_sequence_other_map_sizes_and_shared_frames_vs_oracle(case):
    """Id maps that are not the node's working size (several pixels per target pixel: the last entry wins; fewer: holes keep the
    base draw; a non-integer ratio), and several id frames writing one latent frame (loaders.py:218-243)."""
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    from stable_renderer_b200.loaders import CreateNoiseSequenceFromIdMap
    size, seed = 512, 321
    H, fi = {"map_1024_on_sd15": (1024, [0, 1]), "map_256_on_sd15": (256, [1, 0]), "map_384_on_sd15": (384, [0, 1]),
             "two_id_frames_one_latent_frame": (512, [0, 0, 2])}[case]
    F = len(fi)
    ids = synthetic.make_ids(F, H, H, tex_h=150, tex_w=150, frac_2048=0.1, seed=9)
    n_unique = len(np.unique(O.vertex_screen_info(ids.numpy(), fi)[:, 3]))
    bl, bn, kl, kn = _node_draws(seed, size, n_unique)
    idm = IDMap(tensor=ids.cuda(), frame_indices=fi)
    for option in ("nearest", "max", "mean"):
        want_s, want_n = O.noise_sequence_from_ids(ids.numpy(), fi, bl, bn, kl, kn, size, option)
        out = CreateNoiseSequenceFromIdMap()(idm, seed, "SD15", option, rng_device="cpu")
        if option == "mean":
            np.testing.assert_allclose(out["noise"].cpu().numpy(), want_n, rtol=1e-6, atol=1e-7)
        else:
            assert np.array_equal(out["noise"].cpu().numpy(), want_n), (case, option)
            assert np.array_equal(out["samples"].cpu().numpy(), want_s), (case, option)
    with pytest.raises(IndexError):
        CreateNoiseSequenceFromIdMap()(IDMap(tensor=torch.zeros(1, 64, 128, 4, dtype=torch.int32).cuda()), 1)
