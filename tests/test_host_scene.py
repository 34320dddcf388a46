"""Host front end (mpm_b200/host/*.hpp through libmpm_b200_host.so) against the numpy restatement
of the reference's scene loading / sampling (oracle/scene_oracle.py).  CPU only: no GPU needed."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import scene_oracle as so  # noqa: E402
from mpm_b200 import host  # noqa: E402

SCENES = os.path.join(ROOT, "scenes")
CLI = os.path.join(ROOT, "mpm_b200", "mpm_b200_cli")


def test_options_defaults_forms_and_errors():
    host.lib()  # builds the CLI too
    s = host.Scene("--scene", os.path.join(SCENES, "rubber_duck.toml"), "--N=16", "--particle-count", "10000")
    assert s.n_objects == 2 and len(s.materials) == 1
    host.Scene("--scene", os.path.join(SCENES, "rubber_duck.toml"), "--N", "16", "--particle-count", "2000", "--model", "fixed_corotated",
               "--svd=fast", "--rebin-permille", "50", "--sort-every", "0")
    for bad in (["--bogus", "1"], ["--N"], ["--N", "abc"], ["--N", "-4"], ["positional"], ["--model", "rubber"], ["--svd", "sloppy"], ["-N"],
                ["-N", "x"], ["-M", "16"], ["--sync-every", "0"]):
        r = subprocess.run([CLI] + bad, capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout.strip(), bad  # message + exit(1), like options.h:48-51


def test_grid_size_short_option_like_the_reference():
    """The reference registers the option "N"; cxxopts makes a one-letter name the SHORT option, so its
    README says `-N 16` (README.md:14-16).  All spellings give the same scene."""
    host.lib()
    scene = os.path.join(SCENES, "rubber_duck.toml")
    counts = []
    for form in (["-N", "16"], ["-N16"], ["-N=16"], ["--N", "16"], ["--N=16"]):
        s = host.Scene("--scene", scene, *form, "--particle-count", "10000")
        assert s.N == 16
        counts.append(s.full_count)
    assert len(set(counts)) == 1


@pytest.mark.parametrize("line", ["--scene scenes/rubber_duck.toml -N 16 --particle-count 10000",
                                  "--scene scenes/liquid_bunny.toml -N 32 --particle-count 1000000",
                                  "--scene scenes/snowman.toml -N 64 --particle-count 500000"])
def test_readme_command_lines_parse_verbatim(line):
    """The three example commands of the reference's README (README.md:14-16), argument for argument
    (run from the repository root, like ./docker_run.sh does from the reference's)."""
    host.lib()
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        s = host.Scene(*line.split())
    finally:
        os.chdir(cwd)
    assert s.N == int(line.split()[3]) and s.n_objects >= 1 and s.full_count > 100
    with pytest.raises(host.MpmError):
        host.Scene("--scene", "/nonexistent/scene.toml")


@pytest.mark.parametrize("name", ["snowman", "rubber_duck", "liquid_bunny", "marshmallow_duck"])
def test_scene_files_materials_and_objects(name):
    """TOML subset reader + MaterialModel constructor arithmetic against tomllib + the oracle."""
    import oracle_lib as ol

    path = os.path.join(SCENES, name + ".toml")
    pc = 20000
    s = host.Scene("--scene", path, "--N", "32", "--particle-count", pc)
    doc = so.load_scene_toml(path)
    mats = s.materials
    assert len(mats) == len(doc["material"])
    for got, m in zip(mats, doc["material"]):
        ref = ol.make_material(1.0 / pc, m.get("density", 700.0), m.get("E", 1.4e5), m.get("Nu", 0.2), m.get("hardening", 10.0),
                               m.get("plast_clamp_lower", 0.975), m.get("plast_clamp_higher", 1.0075))
        assert np.array_equal(np.array(got.tolist(), np.float32), ref)
    assert s.n_objects == len(doc["object"])
    names = [m["name"] for m in doc["material"]]
    full = s.full_particles()
    start = 0
    for o, (obj, n, life) in enumerate(zip(doc["object"], s.object_counts(), s.object_lifetimes())):
        part = full[start:start + n]
        start += n
        assert n > 0
        assert (part["material_type"] == names.index(obj["material"])).all()
        assert np.array_equal(part["v"], np.tile(np.float32(obj["velocity"]), (n, 1)))
        assert life[0] == np.float32(obj.get("lifetime_begin", 0.0))
        lo, hi = np.float32(obj["position"]), np.float32(obj["position"]) + np.float32(obj["size"])
        assert (part["x"] >= lo - 1e-6).all() and (part["x"] <= hi + 1e-6).all()  # inside the rescaled bounding box
        assert (part["F"] == np.float32([1, 0, 0, 0, 1, 0, 0, 0, 1])).all() and (part["Jp"] == 1).all() and (part["C"] == 0).all()
    assert all(s.object_substituted())  # this repo ships no .obj files: declared stand-ins
    active = s.active_particles()       # t = 0: objects with lifetime_begin > 0 are not active yet
    n_active = sum(n for n, life in zip(s.object_counts(), s.object_lifetimes()) if life[0] <= 0.0)
    assert len(active) == n_active


def test_reference_scene_files_parse_identically():
    """The reference's own scene files (when the checkout is present) give the same objects."""
    ref_dir = "/root/reference/scenes"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference checkout not present")
    for name in sorted(os.listdir(ref_dir)):
        if not name.endswith(".toml"):
            continue
        ref = host.Scene("--scene", os.path.join(ref_dir, name), "--N", "16", "--particle-count", "3000")
        mine = host.Scene("--scene", os.path.join(SCENES, name), "--N", "16", "--particle-count", "3000")
        assert ref.full_particles().tobytes() == mine.full_particles().tobytes(), name
        assert ref.object_lifetimes() == mine.object_lifetimes()
        assert ref.materials.tobytes() == mine.materials.tobytes()


@pytest.mark.parametrize("mesh,pc", [("sphere.obj", 20000), ("cube.obj", 6000), ("stanford_bunny.obj", 20000)])
def test_sampler_bit_exact_against_oracle(tmp_path, mesh, pc):
    """glibc rand() order, float rounding of the points, winding-number truncation: positions equal
    bit for bit, including the second object (the rand() stream carries over between objects)."""
    scene = tmp_path / "s.toml"
    scene.write_text(f'''
[[material]]
name = "m"
[[object]]
material = "m"
mesh = "{mesh}"
size = 0.37
position = [0.21, 0.13, 0.4]
velocity = [1.0, -2.0, 0.5]
[[object]]
material = "m"
mesh = "{mesh}"
size = 0.2
position = [0.6, 0.6, 0.1]
velocity = [0.0, 0.0, 0.0]
lifetime_begin = 0.01
''')
    s = host.Scene("--scene", str(scene), "--N", "32", "--particle-count", pc, seed=1)
    got = s.full_particles()
    counts = s.object_counts()
    so.srand(1)
    start = 0
    density = int(np.uint32(1.0 / float(np.float32(1.0 / pc))))  # u32(1.0 / material.particleVolume)
    for (size, pos), n in zip(((0.37, [0.21, 0.13, 0.4]), (0.2, [0.6, 0.6, 0.1])), counts):
        V, F, sub = host.load_mesh(str(tmp_path / "meshes" / mesh), np.float32(size), pos)
        assert sub
        Vo = so.rescale(host.load_mesh(str(tmp_path / "meshes" / mesh), 1.0, [0, 0, 0])[0], np.float32(size), pos)
        # rescaling an already rescaled unit mesh is not the same arithmetic; compare the direct path instead
        assert np.abs(V - Vo).max() < 1e-6
        x = so.add_particles(V, F, density)
        assert len(x) == n
        assert np.array_equal(got["x"][start:start + n], x)
        start += n
    assert start == len(got)


def test_winding_number_inside_outside_and_truncation_quirk():
    V, F, _ = host.load_mesh("/nonexistent/meshes/sphere.obj", 1.0, [0.0, 0.0, 0.0])
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1, (4000, 3)).astype(np.float32)
    r = np.linalg.norm(pts - 0.5, axis=1)
    w = host.winding_numbers(V, F, pts)
    assert np.array_equal(w, so.winding_numbers(V, F, pts))  # same float sum, face by face
    assert np.abs(w[r < 0.49] - 1).max() < 1e-5 and np.abs(w[r > 0.51]).max() < 1e-5
    # the reference stores the float winding number in an int matrix (src/mpm.cu:354, 380-383): inside
    # points whose float sum lands just below 1 truncate to 0 and are dropped.  Measured here:
    kept = (w[r < 0.49].astype(np.int32) == 1).mean()
    assert 0.2 < kept <= 1.0
    print(f"winding-number truncation keeps {100 * kept:.1f}% of interior sample points")
    # closed stand-ins: torus (genus 1) and cube
    for name, inside, outside in (("stanford_bunny.obj", [0.5, 0.5 * 0.9 / 2.9, 0.15], [0.5, 0.15, 0.5]), ("cube.obj", [0.5, 0.5, 0.5], [1.2, 0.5, 0.5])):
        V, F, _ = host.load_mesh("/nonexistent/meshes/" + name, 1.0, [0.0, 0.0, 0.0])
        w = host.winding_numbers(V, F, np.float32([inside, outside]))
        assert abs(w[0] - 1) < 1e-5 and abs(w[1]) < 1e-5, name


def test_obj_reader_forms(tmp_path):
    (tmp_path / "meshes").mkdir()
    obj = tmp_path / "meshes" / "tet.obj"
    obj.write_text("# a tetrahedron, mixed corner forms, one quad split off\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\n"
                   "f 1 3 2\nf 1/1 2/1 4/1\nf 2/1/1 3/1/1 4/1/1\nf 1//1 4//1 3//1\n")
    V, F, sub = host.load_mesh(str(obj), 2.0, [1.0, 1.0, 1.0])
    assert not sub and V.shape == (4, 3) and F.shape == (4, 3)
    assert np.allclose(V.min(0), 1.0) and np.allclose(V.max(0), 3.0)  # longest edge -> size, lowest corner -> position
    w = host.winding_numbers(V, F, np.float32([[1.3, 1.3, 1.3], [2.9, 2.9, 2.9]]))
    assert abs(abs(w[0]) - 1) < 1e-5 and abs(w[1]) < 1e-5


def _mesh_checks(V, F):
    edges = np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1)
    _, counts = np.unique(edges, axis=0, return_counts=True)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    volume = np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0
    return counts, volume


def test_marching_tetrahedra_sphere():
    G, r = 40, 13.3
    i, j, k = np.meshgrid(*[np.arange(G, dtype=np.float64)] * 3, indexing="ij")
    S = r - np.sqrt((i - 19.3) ** 2 + (j - 20.1) ** 2 + (k - 18.7) ** 2)
    V, F = host.marching_tetrahedra(S)
    counts, volume = _mesh_checks(V, F)
    assert (counts == 2).all()                                   # closed 2-manifold
    assert abs(volume / (4 / 3 * math.pi * r ** 3) - 1) < 0.01    # outward orientation, right size
    assert np.abs(np.linalg.norm(V - [19.3, 20.1, 18.7], axis=1) - r).max() < 0.1


def test_mesh_builder_and_particle_writer(tmp_path):
    s = host.Scene("--scene", os.path.join(SCENES, "liquid_bunny.toml"), "--N", "32", "--particle-count", "40000", "--mesh-grid", "48",
                   "--mesh-particle-radius", "2")
    nv, nf = s.compute_mesh(tmp_path / "m.obj")
    assert nv > 500 and nf > 1000
    V = np.array([l.split()[1:] for l in open(tmp_path / "m.obj") if l.startswith("v ")], np.float64)
    F = np.array([l.split()[1:] for l in open(tmp_path / "m.obj") if l.startswith("f ")], np.int64) - 1
    assert len(V) == nv and len(F) == nf
    toks = [t for l in open(tmp_path / "m.obj") if l.startswith("v ") for t in l.split()[1:]]
    assert all(t == "%0.17g" % float(t) for t in toks)           # the writer's own formatter prints what %0.17g prints
    counts, volume = _mesh_checks(V, F)
    assert (counts == 2).all() and volume > 0
    x = s.active_particles()["x"]
    assert (V.min(0) < x.min(0)).all() and (V.max(0) > x.max(0)).all()  # the surface wraps the particles (radius 2 voxels)
    assert (V.min(0) > x.min(0) - 3.0 / 48).all() and (V.max(0) < x.max(0) + 3.0 / 48).all()
    s.write_particles(tmp_path / "p.pda")
    lines = open(tmp_path / "p.pda").read().splitlines()
    assert lines[0] == "ATTRIBUTES" and lines[1].split() == ["id", "position", "velocity", "radius"] and lines[3].split() == ["I", "V", "V", "R"]
    assert lines[4] == f"NUMBER_OF_PARTICLES: {len(x)}" and lines[5] == "BEGIN DATA"
    data = np.array([l.split() for l in lines[6:]], np.float64)
    assert np.array_equal(data[:, 0], np.arange(len(x))) and np.array_equal(data[:, 1:4].astype(np.float32), x) and (data[:, 7].astype(np.float32) == np.float32(0.1)).all()


def test_bgeo_particle_file(tmp_path):
    """The reference's particle dump format (src/main.cu:109: particles_%d.bgeo through Partio): Houdini
    classic BGEO v5, big-endian, read back here field by field."""
    import struct

    s = host.Scene("--scene", os.path.join(SCENES, "rubber_duck.toml"), "-N", "16", "--particle-count", "10000")
    p = s.active_particles()
    path = tmp_path / "p.bgeo"
    s.write_particles(path)
    b = open(path, "rb").read()
    assert b[:5] == b"BgeoV"
    version, npts, nprims, npg, nprg, npa, nva, npra, na = struct.unpack(">9i", b[5:41])
    assert (version, npts, nprims, npg, nprg, npa, nva, npra, na) == (5, len(p), 1, 0, 0, 3, 0, 1, 0)
    off = 41
    names = []
    for _ in range(3):
        (ln,) = struct.unpack(">H", b[off:off + 2])
        name = b[off + 2:off + 2 + ln].decode()
        size, typ = struct.unpack(">Hi", b[off + 2 + ln:off + 8 + ln])
        names.append((name, size, typ))
        off += 8 + ln + 4 * size
    assert names == [("id", 1, 1), ("velocity", 3, 5), ("radius", 1, 0)]
    rec = np.frombuffer(b, dtype=np.dtype([("x", ">f4", 4), ("id", ">i4"), ("v", ">f4", 3), ("r", ">f4")]), count=len(p), offset=off)
    assert np.array_equal(rec["x"][:, :3].astype(np.float32), p["x"]) and (rec["x"][:, 3] == 1).all()
    assert np.array_equal(rec["id"], np.arange(len(p))) and np.array_equal(rec["v"].astype(np.float32), p["v"])
    assert (rec["r"].astype(np.float32) == np.float32(0.1)).all()
    off += rec.nbytes
    assert b[off:off + 2 + 9] == struct.pack(">H", 9) + b"generator"
    assert b[-2:] == b"\x00\xff"
    prim = off + 11 + 2 + 4 + 4 + 2 + 4
    key, nv = struct.unpack(">Ii", b[prim:prim + 8])
    assert key == 0x8000 and nv == len(p)


def test_mesh_smoothing_and_decimation(tmp_path):
    """--laplacian_smooth (one implicit cotangent-Laplacian step + the reference's unit-area scaling,
    mesh_builder.h:213-245) and --mesh-face-count (shortest-edge collapse, mesh_builder.h:202-208)."""
    common = ["--scene", os.path.join(SCENES, "liquid_bunny.toml"), "--N", "32", "--particle-count", "40000", "--mesh-grid", "40",
              "--mesh-particle-radius", "2"]

    def load(path):
        V = np.array([l.split()[1:] for l in open(path) if l.startswith("v ")], np.float64)
        F = np.array([l.split()[1:] for l in open(path) if l.startswith("f ")], np.int64) - 1
        return V, F

    def area(V, F):
        return 0.5 * np.linalg.norm(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), axis=1).sum()

    host.Scene(*common).compute_mesh(tmp_path / "raw.obj")
    V0, F0 = load(tmp_path / "raw.obj")
    _, vol0 = _mesh_checks(V0, F0)
    # decimation: face count reached, still a closed 2-manifold, volume kept within a few per cent
    target = len(F0) // 4
    host.Scene(*common, "--mesh-face-count", str(target)).compute_mesh(tmp_path / "dec.obj")
    V1, F1 = load(tmp_path / "dec.obj")
    counts, vol1 = _mesh_checks(V1, F1)
    assert len(F1) <= target and len(F1) > target - 4 and (counts == 2).all()
    assert abs(vol1 / vol0 - 1) < 0.05
    assert len(np.unique(F1)) == len(V1)   # vertices compacted
    # smoothing: same connectivity, unit area afterwards (the reference divides by sqrt(area)), and before
    # that normalisation the surface is smoother: the total area of the rescaled mesh shrinks
    host.Scene(*common, "--laplacian_smooth", "1").compute_mesh(tmp_path / "smooth.obj")
    V2, F2 = load(tmp_path / "smooth.obj")
    assert np.array_equal(F2, F0) and abs(area(V2, F2) - 1.0) < 1e-9
    # undo the normalisation with the raw mesh's scale: centroid-relative sizes compare
    c0, c2 = V0.mean(0), V2.mean(0)
    ratio = np.linalg.norm(V2 - c2, axis=1).mean() / np.linalg.norm(V0 - c0, axis=1).mean()
    V2r = c0 + (V2 - c2) / ratio
    rough = lambda V: np.linalg.norm(V[F0[:, 0]] + V[F0[:, 1]] + V[F0[:, 2]] - 3 * V[F0[:, 0]], axis=1).std()
    assert area(V2r, F0) < area(V0, F0)     # mean-curvature flow shrinks area at equal mean radius
    assert rough(V2r) <= rough(V0) * 1.05


def test_toml_subset_edge_cases(tmp_path):
    """Comments, multi-line arrays, literal strings, escapes, underscores and exponents parse like
    tomllib; constructs outside the subset are rejected instead of being misread."""
    good = tmp_path / "good.toml"
    good.write_text('''# leading comment
[[material]]   # trailing comment
name = 'lit # not a comment'
density = 1_000   # underscore
E = 1.4E+5
Nu = 2e-1
[[material]]
name = "esc \\"quoted\\" \\\\ tab\\t"
[[object]]
material = 'lit # not a comment'
mesh = "cube.obj"
size = 0.25
position = [
  0.1,   # x
  0.2,
  0.3,
]
velocity = [0, -1, 0]
''')
    s = host.Scene("--scene", str(good), "--N", "16", "--particle-count", "4000")
    doc = so.load_scene_toml(str(good))
    assert doc["material"][0]["name"] == "lit # not a comment" and doc["material"][1]["name"] == 'esc "quoted" \\ tab\t'
    mats = s.materials
    assert len(mats) == 2 and s.n_objects == 1
    assert np.isclose(mats[0]["particleMass"], np.float32(1000.0) * np.float32(1.0 / 4000))
    p = s.full_particles()
    assert (p["material_type"] == 0).all() and (p["v"] == np.float32([0, -1, 0])).all()
    assert (p["x"] >= np.float32([0.1, 0.2, 0.3]) - 1e-6).all() and (p["x"] <= np.float32([0.35, 0.45, 0.55]) + 1e-6).all()
    for name, text in (("table", "[material]\nname = 'x'\n"), ("dotted", "[[material]]\na.b = 1\n"), ("unterminated", "[[material]]\nname = \"x\n"),
                       ("garbage", "[[material]]\ndensity = 7 8\n"), ("array", "[[object]]\nposition = [1, 2\n"),
                       ("noposition", "[[material]]\nname='m'\n[[object]]\nmaterial='m'\nvelocity=[0,0,0]\n")):
        bad = tmp_path / (name + ".toml")
        bad.write_text(text)
        with pytest.raises(host.MpmError):
            host.Scene("--scene", str(bad), "--N", "16", "--particle-count", "1000")


def test_obj_negative_indices_and_polygons(tmp_path):
    (tmp_path / "meshes").mkdir()
    obj = tmp_path / "meshes" / "box.obj"
    # a cube from 6 quads, the last two faces with negative (relative) indices
    obj.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 1 1 1\nv 0 1 1\n"
                   "f 1 4 3 2\nf 5 6 7 8\nf 1 2 6 5\nf 4 8 7 3\nf -8 -4 -1 -5\nf -7 -6 -2 -3\n")
    V, F, sub = host.load_mesh(str(obj), 1.0, [0.0, 0.0, 0.0])
    assert not sub and F.shape == (12, 3)
    w = host.winding_numbers(V, F, np.float32([[0.5, 0.5, 0.5], [1.5, 0.5, 0.5]]))
    assert abs(abs(w[0]) - 1) < 1e-5 and abs(w[1]) < 1e-5
    Vs, Fs, _ = host.load_mesh("/nonexistent/meshes/cube.obj", 1.0, [0.0, 0.0, 0.0])
    pts = np.random.default_rng(3).uniform(-0.2, 1.2, (500, 3)).astype(np.float32)
    inside = ((pts > 0.01) & (pts < 0.99)).all(1)
    outside = ((pts < -0.01) | (pts > 1.01)).any(1)
    ws = host.winding_numbers(Vs, Fs, pts)
    assert np.abs(ws[inside] - 1).max() < 1e-5 and np.abs(ws[outside]).max() < 1e-5
