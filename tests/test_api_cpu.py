"""CPU: host-side API surface -- constructor validation, error texts and control flow that the
reference's own tests pin (tests/test_normalizer_template_unit.py,
tests/torch_interface/test_stain_normalizer_transform.py,
tests/torch_interface/test_correctness_against_references.py L201-213).  No kernel runs here."""
from __future__ import annotations

import pytest
import torch

import stainx_b200
from stainx_b200 import HistogramMatching, Macenko, Reinhard, StainNormalizerBase, StainNormalizerTransform
from stainx_b200.sharding import StatReducer, shard_range


def test_public_names_match_the_reference():
    assert set(stainx_b200.__all__) == {"HistogramMatching", "Macenko", "Reinhard", "StainNormalizerBase", "StainNormalizerTransform", "__version__"}
    for cls in (HistogramMatching, Macenko, Reinhard):
        assert issubclass(cls, StainNormalizerBase)
        for method in ("fit", "transform", "fit_transform"):
            assert callable(getattr(cls, method))
    from stainx_b200.backends.torch_cuda_backend import CUDA_AVAILABLE, HistogramMatchingCUDA, MacenkoCUDA, ReinhardCUDA  # noqa: F401

    assert CUDA_AVAILABLE is True


@pytest.mark.parametrize("cls", [Reinhard, Macenko, HistogramMatching])
def test_transform_before_fit_raises(cls):
    with pytest.raises(ValueError, match=r"Must call fit\(\) before transform\(\)"):
        cls(device="cuda", backend="torch_cuda").transform(torch.zeros(1, 3, 8, 8))


def test_backend_validation():
    with pytest.raises(ValueError, match="Unsupported backend"):
        Reinhard(backend="torch")  # no pure-torch backend, no dispatch
    with pytest.raises(ValueError, match="Unsupported backend"):
        HistogramMatching(backend="numpy")
    assert Reinhard(device="cuda").backend == "torch_cuda"
    assert Macenko(device="cuda", backend="torch_cuda").backend == "torch_cuda"


def test_macenko_precision_validation():
    with pytest.raises(ValueError, match="precision='fast' requires backend='torch_cuda'"):
        Macenko(backend="torch", precision="fast")
    with pytest.raises(ValueError, match="precision must be"):
        Macenko(precision="ultra")
    assert Macenko(device="cuda", backend="torch_cuda", precision="fast")._precision == "fast"
    assert Macenko(device="cuda")._precision == "stable"
    assert Macenko(device="cuda").normalize_to_0_1 is False


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_backend_requires_cuda_device():
    n = Reinhard(device="cpu")
    with pytest.raises(ValueError, match="CUDA backend requires CUDA device"):
        n.fit(torch.zeros(1, 3, 8, 8))
    from stainx_b200.backends.torch_cuda_backend import ReinhardCUDA

    with pytest.raises(RuntimeError, match="CUDA is not available"):
        ReinhardCUDA(None)


def test_hm_reference_histogram_validation():
    from stainx_b200.backends.torch_cuda_backend import HistogramMatchingCUDA

    from tests.cpu_ops import cpu_backend

    b = cpu_backend(HistogramMatchingCUDA, "cpu", kernel_layer=object())  # test subclass: skips the device check only
    with pytest.raises(ValueError, match="requires CUDA device"):
        HistogramMatchingCUDA("cpu")  # the product class itself refuses CPU devices
    with pytest.raises(ValueError, match="cannot be empty"):
        b._stack_reference([])
    with pytest.raises(TypeError, match="must be a torch.Tensor"):
        b._stack_reference([1, 2, 3])
    with pytest.raises(ValueError, match="1D with 256 elements"):
        b._stack_reference([torch.zeros(255)])
    with pytest.raises(ValueError, match="1D with 256 elements"):
        b._stack_reference(torch.zeros(2, 100))
    one = torch.rand(256)
    assert torch.equal(b._stack_reference(one), one.unsqueeze(0).repeat(3, 1))
    assert tuple(b._stack_reference([one, one * 2]).shape) == (3, 256)  # padded with channel 0


class TestStainNormalizerTransformValidation:
    def test_modes_and_methods(self):
        with pytest.raises(ValueError, match="Unsupported mode"):
            StainNormalizerTransform("reinhard", mode="online")
        with pytest.raises(ValueError, match="Unknown method"):
            StainNormalizerTransform("vahadane", mode="batch")
        with pytest.raises(ValueError, match="requires a reference tensor"):
            StainNormalizerTransform("reinhard", mode="reference")

    def test_normalize_to_0_1_rules(self):
        with pytest.raises(ValueError, match="only applies to Macenko"):
            StainNormalizerTransform("reinhard", mode="batch", normalize_to_0_1=True)
        assert StainNormalizerTransform("macenko", mode="batch").normalizer.normalize_to_0_1 is True
        assert StainNormalizerTransform("macenko", mode="batch", normalize_to_0_1=False).normalizer.normalize_to_0_1 is False
        pre = Macenko(device="cuda", normalize_to_0_1=False)
        assert StainNormalizerTransform(mode="batch", normalizer=pre).normalizer.normalize_to_0_1 is False  # left alone
        assert StainNormalizerTransform(mode="batch", normalizer=pre, normalize_to_0_1=True).normalizer.normalize_to_0_1 is True
        with pytest.raises(ValueError, match="only applies to Macenko"):
            StainNormalizerTransform(mode="batch", normalizer=Reinhard(device="cuda"), normalize_to_0_1=True)

    def test_layout_rules(self):
        with pytest.raises(ValueError, match="only supported for histogram_matching"):
            StainNormalizerTransform("macenko", mode="batch", channel_axis=-1)
        with pytest.raises(ValueError, match="only supported for histogram_matching"):
            StainNormalizerTransform(mode="batch", normalizer=Reinhard(device="cuda"), channel_axis=3)
        t = StainNormalizerTransform(mode="batch", normalizer=HistogramMatching(device="cuda", channel_axis=-1))
        assert t.channel_axis == -1  # follows the prebuilt normalizer
        with pytest.raises(ValueError, match="conflicts with prebuilt"):
            StainNormalizerTransform(mode="batch", normalizer=HistogramMatching(device="cuda", channel_axis=-1), channel_axis=-3)
        assert StainNormalizerTransform("histogram_matching", mode="batch", channel_axis=3).normalizer.channel_axis == 3

    def test_device_rules(self):
        with pytest.raises(ValueError, match="requires a CUDA device"):
            StainNormalizerTransform("reinhard", mode="batch", device="cpu")
        t = StainNormalizerTransform("reinhard", mode="batch")
        with pytest.raises(ValueError, match="requires CUDA tensors when device=None"):
            t(torch.rand(2, 3, 8, 8))

    def test_shape_checks_run_before_any_device_work(self):
        t = StainNormalizerTransform("macenko", mode="batch")
        with pytest.raises(ValueError, match="Expected NCHW"):
            t(torch.rand(2, 8, 8, 3))
        with pytest.raises(ValueError, match="Expected CHW/NCHW"):
            t(torch.rand(8, 8))
        hm = StainNormalizerTransform("histogram_matching", mode="batch", channel_axis=-1)
        with pytest.raises(ValueError, match="channels-last histogram matching expects"):
            hm(torch.rand(2, 3, 8, 8))

    def test_state_dict_carries_no_fitted_parameters(self):
        t = StainNormalizerTransform("macenko", mode="batch")
        assert t.state_dict() == {}
        assert isinstance(t, torch.nn.Module)


def test_shard_range_partitions_contiguously():
    for n, world in ((64, 8), (10, 4), (3, 8), (0, 2), (512, 8)):
        ranges = [shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_reducer_is_identity_without_a_group():
    r = StatReducer()
    assert not r.enabled and r.world_size == 1 and r.rank == 0
    t = torch.arange(4)
    assert r.sum_(t) is t and r.max_(t) is t and r.min_(t) is t
    assert r.sum_int(7, torch.device("cpu")) == 7


def test_sharded_fit_repeats_exactly_after_a_missed_bracket():
    """Host logic of the deterministic recovery in the sharded pooled fit: when the STATUS region reports a missed bracket
    after the sampled fit, the fit is repeated with the exact coarse pass (hist level 2); a miss after THAT raises."""
    import types

    from stainx_b200._native import StainxNativeError
    from stainx_b200.backends.torch_cuda_backend import MacenkoCUDA
    from tests import cpu_ops
    from tests.helpers import he_tile

    levels, state = [], {"miss_first": 1, "miss_always": False}

    class Flaky(cpu_ops.MacenkoWorkspace):
        def hist(self, images, pooled, stage, level, slot0=0):
            levels.append(level)
            return super().hist(images, pooled, stage, level, slot0)

        def select(self, slot0, count, stage, level):
            super().select(slot0, count, stage, level)
            if stage == 1 and level == 1 and (state["miss_always"] or state["miss_first"] > 0):
                state["miss_first"] -= 1
                self.region("status")[0, 0] = 1  # "rank 0 fell outside its bracket"

    layer = types.SimpleNamespace(**{k: getattr(cpu_ops, k) for k in dir(cpu_ops) if not k.startswith("_")})
    layer.MacenkoWorkspace = Flaky
    ref = torch.cat([he_tile(48, 48, 42), he_tile(48, 48, 7, 1.1)])
    want = cpu_ops.cpu_backend(MacenkoCUDA, "cpu")._pooled_fit_sharded(ref)
    levels.clear()
    he, maxc = cpu_ops.cpu_backend(MacenkoCUDA, "cpu", kernel_layer=layer)._pooled_fit_sharded(ref)
    assert levels == [0, 1, 0, 1, 2, 1, 2, 1], levels  # sampled fit, then the exact repeat
    assert torch.allclose(he, want[0], atol=1e-6) and torch.allclose(maxc, want[1], rtol=1e-6)
    state["miss_always"] = True
    with pytest.raises(StainxNativeError, match="exact bracket"):
        cpu_ops.cpu_backend(MacenkoCUDA, "cpu", kernel_layer=layer)._pooled_fit_sharded(ref)


def test_hm_rejects_non_rgb_with_a_clear_error():
    """the reference's torch backend loops over any channel count; this backend is RGB-only and says so."""
    from stainx_b200.backends.torch_cuda_backend import HistogramMatchingCUDA
    from tests.cpu_ops import cpu_backend

    b = cpu_backend(HistogramMatchingCUDA, "cpu")
    with pytest.raises(ValueError, match="RGB images only"):
        b.compute_reference_counts(torch.zeros((1, 1, 8, 8), dtype=torch.uint8))
    with pytest.raises(ValueError, match="RGB images only"):
        b.transform(torch.zeros((1, 4, 8, 8), dtype=torch.uint8), [torch.ones(256) / 256] * 3)
    with pytest.raises(ValueError, match="4-D batch"):
        b.transform(torch.zeros((3, 8, 8), dtype=torch.uint8), [torch.ones(256) / 256] * 3)


def test_bind_host_thread_is_best_effort():
    """No NVML / no GPU in the build container: the helper changes nothing and says so."""
    import os

    from stainx_b200.ingest import bind_host_thread_to_device

    before = os.sched_getaffinity(0)
    assert bind_host_thread_to_device(0) is None or isinstance(bind_host_thread_to_device(0), list)
    if bind_host_thread_to_device(0) is None:
        assert os.sched_getaffinity(0) == before


def test_macenko_fit_transform_is_fit_then_transform_on_the_cpu_stand_in():
    """Normalizer.fit_transform goes through the backend's fit_transform (one library call on the GPU); with a kernel
    layer that lacks the one-call entry it must fall back to fit + transform with the same result and fitted state."""
    from stainx_b200 import Macenko
    from stainx_b200.backends.torch_cuda_backend import MacenkoCUDA
    from tests import cpu_ops
    from tests.helpers import he_batch

    x = he_batch(3, 40, 48)
    one = Macenko(device="cpu", normalize_to_0_1=True)
    one._backend_impl = cpu_ops.cpu_backend(MacenkoCUDA, "cpu")
    out = one.fit_transform(x)
    assert one._is_fitted and out.dtype == torch.float32 and out.shape == x.shape
    two = Macenko(device="cpu", normalize_to_0_1=True)
    two._backend_impl = cpu_ops.cpu_backend(MacenkoCUDA, "cpu")
    want = two.fit(x).transform(x)
    assert torch.equal(out, want) and torch.equal(one._stain_matrix, two._stain_matrix) and torch.equal(one._target_max_conc, two._target_max_conc)
    with pytest.raises(ValueError, match="NCHW with C=3"):
        one.fit_transform(x[:, :2])
