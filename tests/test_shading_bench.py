"""bench_interface (tools/bench_interface): the oracle's known answers on CPU, CUDA vs oracle on the GPU."""
import numpy as np
import pytest

from oracle import oracle
from rodent_b200 import shading_bench as SB

W = H = 64


def host_mesh(textures, keep):
    def ptr(a):
        a = np.ascontiguousarray(a)
        keep.append(a)
        return a.ctypes.data
    return SB.make_mesh(ptr, SB.quad_arrays(), textures, W, H)


def checker():
    px = np.zeros((H, W, 3), np.float32)
    px[(np.add.outer(np.arange(H), np.arange(W)) % 2) == 1] = (1.0, 0.5, 0.25)
    return px.reshape(-1, 3)


def test_constant_texture_gives_kd_over_pi():
    keep = []
    hits, din, dout = SB.reference_hits(5000)
    for border in (SB.BORDER_CLAMP, SB.BORDER_REPEAT, SB.BORDER_CONSTANT):
        for sampler in (SB.SAMPLER_NEAREST, SB.SAMPLER_BILINEAR):
            kd = (np.tile(np.array((0.1, 0.2, 0.3), np.float32), (W * H, 1)), (0.1, 0.2, 0.3), border, sampler)
            mesh = host_mesh((kd, kd, kd), keep)
            got = oracle.bench_interface(mesh, hits, din, dout)
            want = np.array((0.1, 0.2, 0.3), np.float32) * np.float32(1.0 / np.float32(3.14159265359))
            assert np.allclose(got, want, rtol=2e-7, atol=0), (border, sampler)


def test_texel_centres_return_the_texels():
    """Triangle 0 of the quad maps (u, v) to texcoords (-1 + 2v, 1 - 2u - 2v)... with the repeat border a hit whose
    interpolated coordinate is a texel centre must return that texel (nearest), and the bilinear filter at a texel's
    corner (kx = ky = 0) returns it too."""
    keep = []
    px = checker()
    n = 2000
    rng = np.random.default_rng(1)
    x, y = rng.integers(0, W, n), rng.integers(0, H, n)
    # triangle 0 = vertices (0, 1, 2): texcoord = (1-u-v) t0 + u t1 + v t2 with t0 = (-1, 1), t1 = (-1, -1), t2 = (1, -1)
    # choose the target coordinate inside [0, 1) after the repeat border; the lerp reproduces it up to rounding, so aim at centres
    tu, tv = (x + 0.5) / W, (y + 0.5) / H
    v = (tu + 1.0) / 2.0                      # tex.x = -1 + 2 v
    u = (1.0 - tv) / 2.0 - v                  # tex.y = 1 - 2 u - 2 v
    hits = np.zeros(n, SB.TRI_HIT)
    hits["uv"][:, 0], hits["uv"][:, 1] = u, v
    _, din, dout = SB.reference_hits(n)
    tex = (px, (0, 0, 0), SB.BORDER_REPEAT, SB.SAMPLER_NEAREST)
    got = oracle.bench_interface(host_mesh((tex, tex, tex), keep), hits, din, dout)
    want = px.reshape(H, W, 3)[y, x] * np.float32(1.0 / np.float32(3.14159265359))
    assert np.allclose(got, want, rtol=1e-6, atol=0)


def test_constant_border_outside_the_unit_square():
    keep = []
    hits = np.zeros(4, SB.TRI_HIT)
    hits["uv"] = [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (0.5, 0.25)]         # texcoords (-1,1), (-1,-1), (1,-1), (-0.5,-0.5): all outside
    _, din, dout = SB.reference_hits(4)
    tex = (checker(), (0.5, 1.0, 0.2), SB.BORDER_CONSTANT, SB.SAMPLER_BILINEAR)
    got = oracle.bench_interface(host_mesh((tex, tex, tex), keep), hits, din, dout)
    border = np.array((0.5, 1.0, 0.2), np.float32) * np.float32(1.0 / np.float32(3.14159265359))
    assert np.allclose(got, border, rtol=2e-7)
    hits["uv"] = [(0.1, 0.7)] * 4                                          # texcoord (0.4, -0.6)... still outside in y
    assert np.allclose(oracle.bench_interface(host_mesh((tex, tex, tex), keep), hits, din, dout), border, rtol=2e-7)
    hits["uv"] = [(0.05, 0.6)] * 4                                         # texcoord (0.2, -0.3): outside; (u, v) = (0.1, 0.55): (0.1, -0.3)
    hits["id"] = 1                                                         # triangle 1 = vertices (2, 3, 0): t = (1,-1), (1,1), (-1,1)
    hits["uv"] = [(0.6, 0.3)] * 4                                          # (1-.9)(1,-1) + .6 (1,1) + .3 (-1,1) = (0.4, 0.8): inside
    inside = oracle.bench_interface(host_mesh((tex, tex, tex), keep), hits, din, dout)
    assert not np.allclose(inside, border, rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 100_003])
def test_cuda_matches_oracle_bit_for_bit(n):
    keep = []
    rng = np.random.default_rng(n)
    hits, din, dout = SB.reference_hits(max(n, 1))
    hits, din, dout = hits[:n], din[:n], dout[:n]
    hits["uv"] = rng.uniform(-0.2, 1.2, (n, 2)).astype(np.float32)        # also outside the triangle: borders get exercised
    textures = ((checker(), (0.3, 0.2, 0.1), SB.BORDER_CLAMP, SB.SAMPLER_BILINEAR),
                (checker(), (0.5, 1.0, 0.2), SB.BORDER_CONSTANT, SB.SAMPLER_NEAREST),
                (checker(), (0.0, 0.0, 0.0), SB.BORDER_REPEAT, SB.SAMPLER_BILINEAR))
    for perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):                          # every border / filter mode serves as kd once
        tex = tuple(textures[k] for k in perm)
        want = oracle.bench_interface(host_mesh(tex, keep), hits, din, dout)
        if n == 0:
            continue
        got, _ = SB.run_cuda(SB.quad_arrays(), tex, W, H, hits, din, dout)
        assert got.tobytes() == want.tobytes(), f"perm {perm}: {np.abs(got - want).max()}"


@pytest.mark.gpu
def test_reference_workload_on_gpu():
    """The benchmark's own configuration (bench_interface.cpp:97-183): constant textures, 1 Mi hits."""
    keep = []
    hits, din, dout = SB.reference_hits()
    tex = SB.reference_textures()
    got, seconds = SB.run_cuda(SB.quad_arrays(), tex, 1024, 1024, hits, din, dout, repeat=20)
    want = oracle.bench_interface(host_mesh_big(tex, keep), hits, din, dout)
    assert got.tobytes() == want.tobytes()
    assert seconds > 0


def host_mesh_big(textures, keep):
    def ptr(a):
        a = np.ascontiguousarray(a)
        keep.append(a)
        return a.ctypes.data
    return SB.make_mesh(ptr, SB.quad_arrays(), textures, 1024, 1024)


# ---- bench_shading (tools/bench_shading) ----------------------------------------------------------------
def test_bench_shading_oracle_known_answers():
    s_in, mesh = SB.shading_workload(4096)
    s_out = SB.HostStream(4096)
    SB.call_bench_shading(oracle.bench_shading_fn(), s_in, s_out, mesh, 1)
    assert (s_out.field("depth") == 1).all()
    assert (s_out.field("tmin") == np.float32(0.0001)).all() and (s_out.field("tmax") == np.finfo(np.float32).max).all()
    # the bounced ray starts on the quad (z = 0): org + 1 * dir with org.z = -1, dir.z = 1
    assert np.allclose(s_out.field("org_z"), 0.0, atol=1e-6)
    # (the benchmark's incoming directions are not normalised, so neither are the sampled ones: no length check)
    d = np.stack([s_out.field(n) for n in ("dir_x", "dir_y", "dir_z")], 1)
    assert np.isfinite(d).all() and (np.abs(d).max(axis=1) > 0).all()
    assert (s_out.field("mis") > 0).all() and np.isfinite(s_out.field("contrib_g")).all()
    # geometry 0: kd = ks = (0, 1, 0): no red or blue can come out of it, whatever was sampled
    g0 = slice(0, 1024)
    assert (s_out.field("contrib_r")[g0] == 0).all() and (s_out.field("contrib_b")[g0] == 0).all()
    # the random state advanced and differs from ray to ray
    assert (s_out.field("rnd") != s_in.field("rnd")).all()
    # iterating is idempotent: every iteration recomputes the same outputs from the same inputs
    s_out2 = SB.HostStream(4096)
    SB.call_bench_shading(oracle.bench_shading_fn(), s_in, s_out2, mesh, 3)
    assert s_out2.data.tobytes() == s_out.data.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("num_rays,iters", [(4096, 1), (4096, 7), (8, 2), (1000, 1)])
def test_bench_shading_cuda_matches_oracle(num_rays, iters):
    """CUDA's sinf/cosf differ from libm's in the last bits (they feed the hemisphere samples), everything else follows
    the same fp32 operations.  Measured on a B200: most values bit-identical, 99 % within 1.3e-7 relative, the worst
    7e-5 (one ray whose Phong lobe raises that last-bit difference to the 96th power).  Tolerance: 99 % within 5e-7,
    all within 3e-4; integers and the RNG state exactly."""
    from rodent_b200 import lib
    L = lib.load()
    s_in, mesh = SB.shading_workload(num_rays)
    want, got = SB.HostStream(num_rays), SB.HostStream(num_rays)
    SB.call_bench_shading(oracle.bench_shading_fn(), s_in, want, mesh, iters)
    before = L.rodent_b200_launch_count()
    SB.call_bench_shading(L.b200_bench_shading, s_in, got, mesh, iters)
    assert L.rodent_b200_launch_count() == before + 1
    for name in ("depth", "rnd"):
        assert np.array_equal(got.field(name), want.field(name)), name
    for name in ("org_x", "org_y", "org_z", "tmin", "tmax"):
        assert got.field(name).tobytes() == want.field(name).tobytes(), name
    for name in ("dir_x", "dir_y", "dir_z", "mis", "contrib_r", "contrib_g", "contrib_b"):
        a, b = got.field(name).astype(np.float64), want.field(name).astype(np.float64)
        rel = np.abs(a - b) / (np.abs(b) + 1e-6)
        assert np.quantile(rel, 0.99) < 5e-7 and rel.max() < 3e-4, (name, np.quantile(rel, 0.99), rel.max())
    for name in ("id", "geom_id", "prim_id", "t", "u", "v"):          # not written by the benchmark
        assert not got.field(name).any()
