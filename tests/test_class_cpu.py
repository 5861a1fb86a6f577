"""CPU: the standalone ``deism_b200.core_deism.DEISM`` class up to the first GPU stage — YAML
loader, conflict checks, update_room / update_wall_materials / update_freqs, convex wall
construction — against the fixtures minted from the reference and, when the reference tree is
present (this container only), against the reference's own class stage by stage."""
import numpy as np
import pytest

from conftest import Golden
from deism_b200 import convex, data_loader
from deism_b200.core_deism import DEISM


def _ref():
    try:
        from oracle import ref_import
        if not ref_import.reference_available():
            return None
        ref_import.import_reference()
        return ref_import
    except Exception:
        return None


@pytest.mark.parametrize("mode,roomtype", [("RTF", "shoebox"), ("RIR", "shoebox"), ("RTF", "convex"),
                                           ("RIR", "convex")])
def test_packaged_yaml_matches_reference_defaults(mode, roomtype):
    d = DEISM(mode, roomtype, silent=True)
    p = d.params
    assert p["soundSpeed"] == 343.0 and p["airDensity"] == 1.2
    assert p["sourceOrder"] == 0 and p["receiverOrder"] == 0 and p["ifReceiverNormalize"] == 0   # monopoles
    assert p["DEISM_method"] == "MIX" and p["mixEarlyOrder"] == 2 and p["numParaImages"] == 1000
    assert p["givenMaterials"] == ["impedance"] and p["impedance"] == [18] * 6
    if roomtype == "shoebox":
        assert np.array_equal(p["roomSize"], [4, 3, 2.5]) and p["maxReflOrder"] == 25 and p["angDepFlag"] == 1
        assert p["roomVolumn"] == 30.0 and np.array_equal(p["roomAreas"], [7.5, 7.5, 10, 10, 12, 12])
    else:
        assert p["vertices"].shape == (8, 3) and p["wallCenters"].shape == (6, 3) and p["maxReflOrder"] == 5
        assert abs(p["roomVolumn"] - 36.0) < 1e-12 and p["roomAreas"].shape == (6,)
    if mode == "RTF":
        assert (p["startFreq"], p["freqStep"], p["endFreq"]) == (2, 2, 24000)
    else:
        assert p["sampleRate"] == (48000 if roomtype == "shoebox" else 44100) and p["RIRLength"] == 1
    ref = _ref()
    if ref is None:
        return
    r = ref.make_deism(mode, roomtype).params
    assert set(p) == set(r)
    for key in r:
        if key == "updated_where":
            assert p[key] == r[key]
        elif isinstance(r[key], np.ndarray):
            assert np.array_equal(p[key], r[key]) and p[key].dtype == r[key].dtype, key
        else:
            assert p[key] == r[key] and type(p[key]) is type(r[key]), key


def test_command_line_overrides_and_errors():
    d = DEISM("RTF", "shoebox", silent=True,
              argv=["-nro", "7", "-method", "LC", "-xs", "1", "1", "1", "-fmax", "500", "-adrc", "0", "--run",
                    "--quiet"])
    p = d.params
    assert p["maxReflOrder"] == 7 and p["DEISM_method"] == "LC" and p["endFreq"] == 500 and p["angDepFlag"] == 0
    assert np.array_equal(p["posSource"], [1, 1, 1]) and p["silentMode"] == 1
    with pytest.raises(ValueError):
        DEISM("RTF", "shoebox", silent=True, argv=["-t60", "0.3", "--run"])      # impedance (YAML) and T60
    with pytest.raises(ValueError):
        DEISM("RTF", "sphere", silent=True)
    with pytest.raises(ValueError):
        DEISM("XYZ", "shoebox", silent=True)
    with pytest.raises(ValueError):
        DEISM("RTF", "shoebox", silent=True).update_room(roomDimensions=np.ones((2, 3)))
    c = DEISM("RIR", "convex", silent=True)
    with pytest.raises(ValueError):
        c.update_wall_materials(0.5, None, "reverberationTime")
    with pytest.raises(ValueError):
        data_loader.ConflictChecks.wall_material_checks({"givenMaterials": ["impedance"], "impedance": [1, 2, 3],
                                                         "roomType": "shoebox"})
    with pytest.raises(FileNotFoundError):
        data_loader.readYaml("no_such_config.yml")
    s = DEISM("RTF", "shoebox", silent=True)
    s.params["shoeboxImageCalcVersion"] = "v1"
    with pytest.raises(NotImplementedError):
        s.update_source_receiver()


@pytest.mark.parametrize("name,mode,materials,overrides", [
    ("c1_full", "RTF", None, {}),
    ("c2_full_mono", "RIR", (1.0 / 3.0, None, "reverberationTime"), {"sampleRate": 48000}),
    ("c5_full_N5", "RIR", (2.0 / 3.0, None, "reverberationTime"), {"sampleRate": 48000}),
    ("s5_shoebox_rtf_lc_o7_mono_noang", "RTF", None, {"endFreq": 300, "angDepFlag": 0}),
])
def test_shoebox_host_stages_equal_fixture(name, mode, materials, overrides):
    g = Golden(name)
    gp = g.group("params")
    d = DEISM(mode, "shoebox", silent=True)
    d.params.update(overrides)
    if name.startswith("s5"):
        d.update_room(roomDimensions=np.array([5.3, 3.1, 2.7]))
    d.update_wall_materials(*(materials or ()))
    d.update_freqs()
    p = d.params
    assert p["reverberationTime"] == float(gp["reverberationTime"])
    assert np.array_equal(p["freqs"], gp["freqs"]) and np.array_equal(p["waveNumbers"], gp["waveNumbers"])
    assert np.array_equal(np.asarray(p["impedance"])[:, 0], gp["impedance_col0"])
    assert np.asarray(p["impedance"]).shape == (6, len(gp["freqs"]))
    assert all(key in p for key in ("n1", "n2", "n3", "freqs_bands", "absorption"))
    if mode == "RIR":
        assert p["nSamples"] == int(p["sampleRate"] * p["RIRLength"])


@pytest.mark.parametrize("name,sample_rate", [("c4_full", None), ("x1_convex_rir_mix_o4_mono", 4000)])
def test_convex_host_stages_and_walls_equal_fixture(name, sample_rate):
    g = Golden(name)
    gp = g.group("params")
    d = DEISM("RIR", "convex", silent=True)
    d.params["maxReflOrder"] = int(gp["maxReflOrder"])
    if sample_rate:
        d.params["sampleRate"] = sample_rate
    d.update_wall_materials()
    d.update_freqs()
    p = d.params
    assert p["reverberationTime"] == float(gp["reverberationTime"])
    assert np.array_equal(p["freqs"], gp["freqs"])
    walls = d.room_convex.walls
    assert len(walls) == int(g.z["walls/n"])
    for i, w in enumerate(walls):
        for key in ("normal", "origin", "corners", "reflection_matrix"):
            assert np.array_equal(getattr(w, key), g.z[f"walls/{i}/{key}"].reshape(getattr(w, key).shape)), (i, key)
        assert np.array_equal(np.asarray(w.impedence_bands), g.z[f"walls/{i}/impedance"])
        # the plane basis is the reference's up to an isometry of the plane: orthonormal, normal to n,
        # and the projected polygon has the reference's (positive) orientation and edge lengths
        assert np.allclose(w.basis.T @ w.basis, np.eye(2), atol=1e-6) and np.allclose(w.normal @ w.basis, 0, atol=1e-6)
        ref_flat = g.z[f"walls/{i}/flat_corners"]
        assert convex._area_2d(w.flat_corners) > 0
        assert np.isclose(convex._area_2d(w.flat_corners), convex._area_2d(ref_flat), rtol=1e-5)
    # the reflection tree over these walls is the reference's, bit for bit (CPU restatement of the engine)
    from oracle import oracle

    out = oracle.convex_images(convex.pack_walls(walls), p["posSource"], p["posReceiver"], int(p["maxReflOrder"]))
    e = g.group("engine")
    for key in ("sources", "attenuations", "reflection_matrix", "orders", "gen_walls"):
        assert np.array_equal(out[key], e[key]), key


def test_random_convex_rooms_against_compiled_reference_walls():
    """Wall_deism (float32 mirror of wall.cpp:338-397) vs the compiled reference on sheared /
    tilted hexahedra: normal, origin, corner order and reflection matrix bit-equal."""
    try:
        from oracle import ref_import
        lr = ref_import.load_libroom()
    except Exception:
        pytest.skip("oracle/_ref/libroom_deism not built")
    rng = np.random.default_rng(11)
    base = np.array([[0, 0, 0], [0, 0, 3.5], [0, 3, 2.5], [0, 3, 0], [4, 0, 0], [4, 0, 3.5], [4, 3, 2.5], [4, 3, 0]], float)
    for trial in range(25):
        A = np.eye(3) + 0.15 * rng.standard_normal((3, 3))             # affine image stays a convex hexahedron
        V = base @ A.T + rng.uniform(-1, 1, 3)
        hull, groups = convex._faces_by_normal(V)
        cen = V.mean(axis=0)
        for facets in groups:
            pts = np.unique(np.concatenate([hull.points[hull.simplices[i]] for i in facets]), axis=0)
            try:
                mine = convex.Wall_deism(pts, cen, np.array([18.0]))
                mine_err = None
            except RuntimeError as exc:
                mine, mine_err = None, exc
            z = np.array([18.0], np.float32)
            try:
                ref = lr.Wall_deism(pts.T, cen.T, z, np.zeros(1, np.float32), np.zeros(1, np.float32), "w")
            except RuntimeError:
                assert mine_err is not None
                continue
            assert mine_err is None
            assert np.array_equal(mine.normal, np.asarray(ref.normal).reshape(3))
            assert np.array_equal(mine.origin, np.asarray(ref.origin).reshape(3))
            assert np.array_equal(mine.corners, np.asarray(ref.corners))
            assert np.array_equal(mine.reflection_matrix, np.asarray(ref.reflection_matrix))


def test_yaml_lookup_order_and_legacy_keys(tmp_path, monkeypatch):
    """readYaml: ./examples/<name> wins over ./<name> wins over the packaged default (data_loader.py:398-409);
    the old YAML key `absorpCoefficient` is still read (data_loader.py:843-848); convex flags parse."""
    import yaml

    base = yaml.safe_load(open(f"{data_loader.CONFIG_DIR}/configSingleParam_RTF.yml"))
    monkeypatch.chdir(tmp_path)
    assert DEISM("RTF", "shoebox", silent=True).params["maxReflOrder"] == 25             # packaged default
    cfg = dict(base)
    cfg["Reflections"] = {"angleDependentFlag": 0, "absorpCoefficient": [0.1] * 6, "maxReflectionOrder": 3}
    (tmp_path / "configSingleParam_RTF.yml").write_text(yaml.safe_dump(cfg, sort_keys=False))
    p = DEISM("RTF", "shoebox", silent=True).params
    assert p["maxReflOrder"] == 3 and p["givenMaterials"] == ["absorption"] and p["absorption"] == [0.1] * 6
    (tmp_path / "examples").mkdir()
    cfg["Reflections"]["maxReflectionOrder"] = 7
    (tmp_path / "examples" / "configSingleParam_RTF.yml").write_text(yaml.safe_dump(cfg, sort_keys=False))
    assert DEISM("RTF", "shoebox", silent=True).params["maxReflOrder"] == 7
    c = DEISM("RIR", "convex", silent=True, argv=["-ifconvex", "1", "-nro", "3", "-fs", "8000", "--run", "--quiet"])
    assert c.params["convexRoom"] == 1 and c.params["maxReflOrder"] == 3 and c.params["sampleRate"] == 8000
    with pytest.raises(SystemExit):
        DEISM("RIR", "convex", silent=True, argv=["-room", "1", "2", "3"])              # shoebox-only flag


def test_material_no_order_means_minus_one(tmp_path, monkeypatch):
    import yaml

    cfg = yaml.safe_load(open(f"{data_loader.CONFIG_DIR}/configSingleParam_RTF.yml"))
    cfg["Reflections"]["maxReflectionOrder"] = None
    monkeypatch.chdir(tmp_path)
    (tmp_path / "configSingleParam_RTF.yml").write_text(yaml.safe_dump(cfg, sort_keys=False))
    assert DEISM("RTF", "shoebox", silent=True).params["maxReflOrder"] == -1              # data_loader.py:823-829
