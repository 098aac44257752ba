"""CPU tests of the host logic added in round 2 (no GPU, no native compute calls): allocation buckets, the model-gradient
arena layout of the fused projection, activation flags, the densification buffer pools, the field-of-view cache and the
optimizer's state surface."""
import types

import pytest
import torch


def test_row_capacity_buckets():
    from gaussianhaircut_b200._alloc import row_capacity, empty_rows
    prev = 0
    for P in list(range(0, 9000, 37)) + [100_000, 500_000, 500_001, 562_500, 2_000_000, 2_003_811, 3_070_000]:
        cap = row_capacity(P)
        assert cap >= P and cap >= 4096
        assert cap <= P + max(4096, P // 8 + 1), (P, cap)              # one bucket at most: 4096 rows or 12.5 %
        if P > 4096:
            assert cap % 4096 == 0
        assert cap >= prev or P < 9000
        prev = cap if P >= 9000 else prev
    # a model that grows by a fraction of a percent stays in its bucket: the allocator's cached blocks are reused
    assert row_capacity(2_000_847) == row_capacity(2_002_158) == row_capacity(2_003_811)
    t = empty_rows(5000, (3,), torch.float32, torch.device("cpu"))
    assert t.shape == (5000, 3) and t.is_contiguous() and t.untyped_storage().nbytes() >= row_capacity(5000) * 12


def test_projection_gradient_arena_layout():
    """61 floats per Gaussian in ONE buffer, every segment padded to 4 floats (16-byte aligned views, a 4-float-granular
    all-reduce never spills from one segment into the next), views shaped like the model's parameters."""
    from gaussianhaircut_b200 import projection as pj
    for P in (1, 2, 3, 7, 50_001):
        n = pj.grad_arena_floats(P)
        assert n % 4 == 0 and n >= 61 * P and n <= 61 * P + 4 * 8
        assert pj.grad_arena_floats(P, with_dirs=True) == n + (3 * P + 3) // 4 * 4
        storage = torch.zeros(n + 5)
        views = pj.carve_grad_arena(storage, P)
        assert set(views) == {"rotation", "xyz", "scaling", "f_dc", "f_rest", "opacity", "label", "conf"}
        assert views["f_dc"].shape == (P, 1, 3) and views["f_rest"].shape == (P, 15, 3) and views["rotation"].shape == (P, 4)
        seen = torch.zeros(n + 5, dtype=torch.int32)
        for k, v in views.items():
            assert v.data_ptr() % 16 == 0, k
            off = (v.data_ptr() - storage.data_ptr()) // 4
            seen[off:off + v.numel()] += 1
            v.fill_(1.0)
        assert int(seen.max()) == 1 and int(seen.sum()) == 61 * P          # disjoint, nothing missing
        assert float(storage.sum()) == 61 * P
    with pytest.raises(RuntimeError, match="gradient arena"):
        pj.carve_grad_arena(torch.zeros(10), 5)
    with pytest.raises(RuntimeError, match="gradient arena"):
        pj.carve_grad_arena(torch.zeros(1000, dtype=torch.float64), 5)


def test_projection_flags_and_cpu_rejection():
    from gaussianhaircut_b200 import projection as pj
    # bits 0-1 scale, 2-3 opacity, 4-5 label, 6-7 confidence, 8-9 direction (include/gh_rasterizer.h)
    assert pj.encode_flags(pj.GAUSSIAN_MODEL) == (1 | (1 << 2) | (1 << 4) | (1 << 6) | (0 << 8))
    assert pj.encode_flags(pj.HAIR_MODEL) == (0 | (2 << 2) | (2 << 4) | (1 << 6) | (1 << 8))
    f = pj.encode_flags(pj.HEAD_PRECOMP)
    assert (f >> 4) & 3 == 3 and (f >> 6) & 3 == 3 and (f >> 8) & 3 == 2     # label 0, confidence 0, no direction
    z = torch.zeros
    with pytest.raises(RuntimeError, match="no CPU path"):
        pj.pack_inputs(z(4, 3), z(4, 3), z(4, 4), None, z(4, 1, 3), z(4, 15, 3), z(4, 1), z(4, 1), z(4, 1), z(4, 4), z(4, 4), z(3),
                       0.5, 0.5, 64, 64, 3, 1.0, pj.GAUSSIAN_MODEL)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        pj.pack_inputs(z(4, 2), z(4, 3), z(4, 4), None, z(4, 1, 3), z(4, 15, 3), z(4, 1), z(4, 1), z(4, 1), z(4, 4), z(4, 4), z(3),
                       0.5, 0.5, 64, 64, 3, 1.0, pj.GAUSSIAN_MODEL)


def test_densify_pools_alternate_and_grow():
    """Source and destination of a compaction alternate between two pooled buffers per tensor; a request that fits is
    served without a new allocation, one that does not replaces only its own slot."""
    from gaussianhaircut_b200.densify import _pool_alloc
    g = types.SimpleNamespace()
    dev = torch.device("cpu")
    a = _pool_alloc(g, "xyz", 1000, (3,), dev, None)
    assert a.shape == (1000, 3) and a.is_contiguous()
    b = _pool_alloc(g, "xyz", 1100, (3,), dev, a)                  # `a` is the live tensor: must not be handed out again
    assert b.data_ptr() != a.data_ptr() and b.shape == (1100, 3)
    c = _pool_alloc(g, "xyz", 1200, (3,), dev, b)                  # fits slot 0 thanks to the 25 % headroom
    assert c.data_ptr() == a.data_ptr()
    d = _pool_alloc(g, "xyz", 1250, (3,), dev, c)                  # fits slot 1 as well
    assert d.data_ptr() == b.data_ptr()
    e = _pool_alloc(g, "xyz", 5000, (3,), dev, d)                  # outgrows slot 0: re-allocated, slot 1 untouched
    assert e.shape == (5000, 3) and g._gh_pools["xyz"][1].data_ptr() == b.data_ptr()
    f = _pool_alloc(g, "opacity", 1000, (1,), dev, None)           # pools are per tensor name
    assert f.data_ptr() not in (e.data_ptr(), d.data_ptr()) and set(g._gh_pools) == {"xyz", "opacity"}


def test_tan_half_cache():
    """renderer._tan_half: one device read per field-of-view tensor, invalidated by in-place updates, never cached for a
    trainable field of view."""
    import math
    from gaussianhaircut_b200 import renderer
    assert renderer._tan_half(1.0) == math.tan(0.5)
    fov = torch.tensor(0.8)
    v0 = renderer._tan_half(fov)
    assert abs(v0 - math.tan(0.4)) < 1e-6
    assert renderer._TAN_CACHE[id(fov)][2] == v0 and renderer._tan_half(fov) == v0
    fov.mul_(0.5)                                                   # in-place change bumps the version counter
    assert abs(renderer._tan_half(fov) - math.tan(0.2)) < 1e-6
    trainable = torch.tensor(0.8, requires_grad=True)
    renderer._tan_half(trainable)
    assert id(trainable) not in renderer._TAN_CACHE or renderer._TAN_CACHE[id(trainable)][0]() is not trainable


def test_fused_adam_has_no_cpu_path():
    """FusedAdam refuses CPU parameters at construction (there is no fallback optimizer)."""
    from gaussianhaircut_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(8, 3))
    with pytest.raises(ValueError, match="CUDA"):
        FusedAdam([{"params": [p], "lr": 1e-3, "name": "xyz"}], eps=1e-15)
