"""CPU: host-side planning logic of the C ABI (no kernel is launched): tcgen05 tile plans, packed-weight image sizes,
split-K workspaces and the dispatch rules added for thin / grouped / waveform-input layers."""
import ctypes

import pytest

from kantts_b200 import _lib, ops
from kantts_b200._lib import KT_PATH_AUTO, KT_PATH_FFMA, KT_PATH_TC


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _desc(B, t_in, nsub=1, **kw):
    path = kw.pop("path", KT_PATH_AUTO)
    spec = ops.ConvSpec(path=path, **kw)
    return spec, spec.desc(B, nsub, t_in)


def test_dense_layer_tiles(lib):
    # 1024 -> 1024 k5 (period discriminator): N tiles of 256, 16 K chunks of 64
    _, d = _desc(32, 34, nsub=3, c_in=1024, c_out=1024, kernel=5, pad_left=2, pad_right=2)
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) == 256 and lib.kt_conv1d_tc_plan(ctypes.byref(d), 1) == 256
    # image = taps x K chunks x N tiles x (hi + lo) x NT x 64 bf16
    assert lib.kt_conv1d_tc_image_bytes(ctypes.byref(d), 0) == 5 * 16 * 4 * 2 * 256 * 64 * 2
    # generator resblock conv: one N tile of the layer's width
    _, d = _desc(16, 8192, c_in=32, c_out=32, kernel=7, pad_left=6)
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) == 32


def test_thin_groups_share_block_diagonal_tiles(lib):
    # 128 -> 256, 16 groups (8 -> 16 channels per group): 8 groups per tile = K 64 x N 128, 2 N tiles
    _, d = _desc(16, 2048, c_in=128, c_out=256, kernel=41, stride=4, pad_left=20, pad_right=20, groups=16)
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) == 128
    assert lib.kt_conv1d_tc_image_bytes(ctypes.byref(d), 0) == 41 * 1 * 2 * 2 * 128 * 64 * 2
    # its data gradient contracts 16 channels per group and produces 8: 4 groups per tile = K 64 x N 32
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 1) == 32
    # 64-channel groups already fill a K chunk: one group per tile
    _, d = _desc(16, 32, c_in=1024, c_out=1024, kernel=41, pad_left=20, pad_right=20, groups=16)
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) == 64


def test_waveform_input_layers_leave_the_tensor_path(lib):
    kw = dict(c_in=1, c_out=32, kernel=5, stride=3, pad_left=2, pad_right=2, act_out=1, act_out_slope=0.1)
    _, d = _desc(32, 2731, nsub=3, **kw)
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) == 0            # forward: FIR kernel (thin.cu)
    assert lib.kt_conv1d_bwd_weight_tc_workspace(ctypes.byref(d)) == 0
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 1) > 0             # its data gradient stays on tcgen05
    _, d = _desc(32, 2731, nsub=3, path=KT_PATH_TC, **kw)            # explicitly requested: still available
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) > 0 and lib.kt_conv1d_bwd_weight_tc_workspace(ctypes.byref(d)) > 0


def test_upsampled_conv_data_gradient_plan(lib):
    spec, d = _desc(16, 32, c_in=512, c_out=256, kernel=7, pad_left=6, upsample=8, act_in=1, act_in_slope=0.1)
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 0) > 0
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d), 1) == 0            # not directly ...
    s2 = spec.without_upsample()
    d2 = s2.desc(16, 1, 32 * 8)
    assert d2.t_out == d.t_out and s2.act_in == 0 and s2.upsample == 1
    assert lib.kt_conv1d_tc_plan(ctypes.byref(d2), 1) > 0            # ... but as the plain conv over the up-sampled rows


def test_split_k_workspace_is_whole_slices_of_the_gradient(lib):
    for kw, B, T, nsub in ((dict(c_in=1024, c_out=1024, kernel=5, pad_left=2, pad_right=2), 32, 34, 3),
                           (dict(c_in=32, c_out=32, kernel=11, pad_left=10), 16, 8192, 1),
                           (dict(c_in=128, c_out=256, kernel=41, stride=4, pad_left=20, pad_right=20, groups=16), 16, 2048, 1)):
        spec, d = _desc(B, T, nsub=nsub, **kw)
        ws = lib.kt_conv1d_bwd_weight_tc_workspace(ctypes.byref(d))
        # one slice per split: the weight gradient + (when the bias gradient rides along as an all-ones MMA unit) the
        # bias gradient padded to a multiple of 4 floats
        per = spec.w_numel + ((spec.c_out + 3) & ~3)
        assert ws > 0 and (ws % spec.w_numel == 0 or ws % per == 0)
        nsplit = ws // per if ws % per == 0 else ws // spec.w_numel
        assert 1 <= nsplit <= 296


def test_grad_items_context_restores_state():
    assert ops._grad_items is None
    with ops.grad_items(4):
        assert ops._grad_items == 4
        with ops.grad_items(2):
            assert ops._grad_items == 2
        assert ops._grad_items == 4
    assert ops._grad_items is None


def test_ffma_path_has_no_tensor_plan(lib):
    _, d = _desc(2, 100, c_in=64, c_out=64, kernel=3, pad_left=2, path=KT_PATH_FFMA)
    spec = ops.ConvSpec(c_in=64, c_out=64, kernel=3, pad_left=2, path=KT_PATH_FFMA)
    assert ops._tc_tile(lib, spec, d, 0) == 0 and ops._wgrad_tc_workspace(lib, spec, d) == 0
