"""-m gpu: bit-exact keying (IDMap.create_vertex_screen_info) and tensor_group_by_then_average on the GPU."""
import numpy as np
import pytest
import torch

import srx_oracle as O
from helpers import assert_close, t2n

pytestmark = pytest.mark.gpu

CASES = ["step_sq64_r8", "step_sq96_r8_perm", "step_sq100_nonint", "step_sq60_to_16", "step_dupframe"]


@pytest.mark.parametrize("name", CASES)
def test_vertex_screen_info_bit_exact_vs_reference(golden, name):
    from stable_renderer_b200.corrmap import IDMap
    g = golden(name)
    idm = IDMap(tensor=torch.from_numpy(g["ids"]).cuda(), frame_indices=[int(v) for v in g["frame_indices"]])
    vsi = idm.create_vertex_screen_info()
    assert vsi.dtype == torch.float32 and tuple(vsi.shape) == g["vsi"].shape
    assert np.array_equal(vsi.cpu().numpy().view(np.uint32), g["vsi"].view(np.uint32))      # all 7 columns, every bit
    assert np.array_equal(t2n(idm.masks), g["masks"])
    assert idm.create_vertex_screen_info() is vsi                                              # cached like corrmap.py:226,278


def test_vertex_screen_info_int16_and_large(golden):
    from stable_renderer_b200 import synthetic
    from stable_renderer_b200.corrmap import IDMap
    g = golden("step_sphere_crop_int16")
    idm = IDMap(tensor=torch.from_numpy(g["ids"]).cuda())
    want = O.vertex_screen_info(g["ids"])
    got = idm.create_vertex_screen_info().cpu().numpy()
    assert got.shape[0] == int(g["n_entries"])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    ids = synthetic.make_ids(5, 200, 200, tex_h=300, tex_w=300, frac_2048=0.1, n_obj=3, seed=5)   # ragged tile counts
    want = O.vertex_screen_info(ids.numpy(), [4, 9, 2, 7, 11])
    got = IDMap(tensor=ids.cuda(), frame_indices=[4, 9, 2, 7, 11]).create_vertex_screen_info().cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_vertex_screen_info_empty():
    from stable_renderer_b200.corrmap import IDMap
    vsi = IDMap(tensor=torch.zeros(2, 16, 16, 4, dtype=torch.int32).cuda()).create_vertex_screen_info()
    assert tuple(vsi.shape) == (0, 7)


def test_group_by_then_average_kats(golden):
    # the reference's docstring examples (source/common_utils/math_utils.py:110-128)
    from stable_renderer_b200.math_utils import tensor_group_by_then_average
    t = torch.tensor([[2, 1, 4], [2, 9, 12], [6, 4, 4], [7, 3, 99], [8, 1, 3]]).cuda()
    (a0,) = tensor_group_by_then_average(t, index_column=0, value_columns=[1, 2])
    assert torch.equal(a0.cpu(), torch.tensor([[5., 8.], [5., 8.], [4., 4.], [3., 99.], [1., 3.]]))
    a1, u1 = tensor_group_by_then_average(t, index_column=1, value_columns=[0], return_unique=True)
    assert torch.equal(a1.cpu(), torch.tensor([[5.], [2.], [6.], [7.], [5.]]))
    assert u1.cpu().tolist() == [1, 3, 4, 9]
    g = golden("group_by_average")
    b, ub = tensor_group_by_then_average(torch.from_numpy(g["big"]).cuda(), -1, [0, 1, 2, 3], return_unique=True)
    assert np.array_equal(t2n(ub), g["ub"])
    assert_close(t2n(b), g["b"], 1e-5, 1e-6)


def test_group_by_then_average_errors():
    from stable_renderer_b200.math_utils import tensor_group_by_then_average
    t = torch.zeros(3, 3).cuda()
    with pytest.raises(ValueError):
        tensor_group_by_then_average(t, 3, [0])
    with pytest.raises(ValueError):
        tensor_group_by_then_average(t, 0, [5])


def test_adain_helper(golden):
    from stable_renderer_b200.math_utils import adaptive_instance_normalization
    g = golden("group_by_average")
    out = adaptive_instance_normalization(torch.from_numpy(g["content"]).cuda(), torch.from_numpy(g["style"]).cuda())
    assert_close(t2n(out), g["adain"], 1e-5, 2e-6)
