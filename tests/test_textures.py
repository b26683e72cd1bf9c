"""Textured materials (MTL map_Kd / map_Ks / map_Ke): the PNG loader against src/driver/image.cpp's behaviour, the binding of
images to materials (converter.cpp:595-602, 748-768, 876-903), the oracle's texturing, and CUDA against the oracle."""
import numpy as np
import pytest
from PIL import Image

from oracle import oracle
from rodent_b200 import render as R

from png_writer import expected_pixels, gamma_lut, write_png

OBJ = """mtllib tex.mtl
v -1 0 1
v 1 0 1
v 1 0 -1
v -1 0 -1
v -1 2 -1
v 1 2 -1
v -0.3 1.9 -0.3
v 0.3 1.9 -0.3
v 0.3 1.9 0.3
v -0.3 1.9 0.3
vt 0 0
vt 3 0
vt 3 3
vt 0 3
vt -0.5 -0.25
vt 1.5 -0.25
vt 1.5 1.75
vt -0.5 1.75
usemtl floor
f 1/1 2/2 3/3 4/4
usemtl wall
f 4/5 3/6 6/7 5/8
usemtl lamp
f 7 8 9 10
"""
MTL = """newmtl floor
Kd 0.5 0.5 0.5
map_Kd {floor}
illum 2
newmtl wall
Kd {wall_kd}
Ks 0.3 0.3 0.3
Ns 20
map_Ks {wall}
illum 2
newmtl lamp
Kd 0.7 0.7 0.7
Ke 12 12 10
illum 2
"""


def checker(n=16, cell=2, seed=3):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
    dark = (np.add.outer(np.arange(n) // cell, np.arange(n) // cell) % 2) == 0
    img[dark, :3] //= 4
    img[..., 3] = 255
    return img


def write_scene(tmp_path, floor="floor.png", wall="wall.png", wall_kd="0.4 0.3 0.2"):
    (tmp_path / "tex.obj").write_text(OBJ)
    (tmp_path / "tex.mtl").write_text(MTL.format(floor=floor, wall=wall, wall_kd=wall_kd))
    return tmp_path / "tex.obj"


def camera(W, H):
    return R.camera((0, 1, 2.5), (0, 0, -1), (0, 1, 0), 60.0, W, H)


# ---- the PNG loader -------------------------------------------------------------------------------------------
def rgba_of(samples, color, depth, palette=None):
    """What libpng hands load_png for these samples: 8-bit RGBA (image.cpp:63-80), alpha 255 without tRNS."""
    s = samples.astype(np.int64)
    if depth == 16:
        s = s >> 8                                                       # png_set_strip_16
    elif depth < 8 and color == 0:
        s = s * 255 // ((1 << depth) - 1)                                # grey expanded to 8 bit
    h, w, _ = s.shape
    out = np.full((h, w, 4), 255, np.uint8)
    if color == 0:
        out[..., :3] = s[..., :1]
    elif color == 2:
        out[..., :3] = s
    elif color == 3:
        out[..., :3] = np.asarray(palette, np.uint8)[samples[..., 0]]
    elif color == 4:
        out[..., :3] = s[..., :1]
        out[..., 3] = s[..., 1]
    else:
        out[...] = s
    return out


VARIANTS = [  # colour type, bit depth, interlaced, width, height
    (2, 8, False, 13, 7), (6, 8, False, 8, 8), (0, 8, False, 5, 9), (4, 8, False, 6, 4), (3, 8, False, 11, 5),
    (3, 4, False, 9, 6), (3, 2, False, 7, 3), (3, 1, False, 17, 2), (0, 1, False, 10, 3), (0, 2, False, 5, 5), (0, 4, False, 7, 7),
    (2, 16, False, 4, 6), (6, 16, False, 3, 3), (0, 16, False, 5, 2), (4, 16, False, 2, 5),
    (2, 8, True, 13, 11), (6, 8, True, 1, 1), (3, 4, True, 9, 9), (0, 1, True, 20, 3), (2, 16, True, 3, 2), (2, 8, True, 2, 1),
]


@pytest.mark.parametrize("color,depth,interlace,w,h", VARIANTS)
def test_png_loader(tmp_path, color, depth, interlace, w, h):
    rng = np.random.default_rng(color * 100 + depth + w)
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color]
    samples = rng.integers(0, 1 << depth, (h, w, ch))
    palette = rng.integers(0, 256, (1 << depth, 3)) if color == 3 else None
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    for first_filter in range(5):                                           # every row filter lands on every row once
        path = tmp_path / f"t{first_filter}.png"
        write_png(path, samples, color, depth, interlace=interlace, first_filter=first_filter, palette=palette,
                  idat_pieces=1 + first_filter % 3)
        tid = scene.add_png(path)
        tex = scene.array("textures")[tid - 1]
        assert (tex["width"], tex["height"]) == (w, h)
        got = scene.array("texture_pixels")[tex["offset"]:tex["offset"] + w * h].reshape(h, w)
        want = expected_pixels(rgba_of(samples, color, depth, palette))
        assert np.array_equal(got & 0xFFFFFF, want & 0xFFFFFF)
        assert np.array_equal(got >> 24, want >> 24)
        if depth == 8 and color in (0, 2, 4, 6) and not interlace:          # PIL reads our file the same way
            pil = np.array(Image.open(path).convert("RGBA"))
            assert np.array_equal(expected_pixels(pil), got)


def test_png_written_by_pil_and_gamma(tmp_path):
    """A file from an independent encoder (PIL picks its own filters), paletted with transparency, and the gamma table:
    0 -> 0, 255 -> 255, 128 -> pow(128/255, 2.2) * 255 = 55.9 -> 55 (truncated, image.cpp:16)."""
    lut = gamma_lut()
    assert (lut[0], lut[255], lut[128], lut[1]) == (0, 255, 55, 0) and (np.diff(lut.astype(int)) >= 0).all()
    img = checker(32)
    Image.fromarray(img).save(tmp_path / "rgba.png")
    Image.fromarray(img[..., :3]).convert("P", palette=Image.ADAPTIVE, colors=16).save(tmp_path / "pal.png", transparency=3)
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    for name in ("rgba.png", "pal.png"):
        tid = scene.add_png(tmp_path / name)
        tex = scene.array("textures")[tid - 1]
        got = scene.array("texture_pixels")[tex["offset"]:tex["offset"] + 32 * 32].reshape(32, 32)
        assert np.array_equal(got, expected_pixels(np.array(Image.open(tmp_path / name).convert("RGBA")))), name
    assert (got >> 24 == 0).any() and (got >> 24 == 255).any()             # tRNS became alpha (image.cpp:75-76)


def test_png_loader_rejects_bad_files(tmp_path, capfd):
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    good = tmp_path / "good.png"
    write_png(good, np.zeros((4, 4, 3), np.int64), 2, 8)
    blob = good.read_bytes()
    k = blob.index(b"IDAT") + 6                                             # a byte of the compressed stream: the chunk's CRC fails
    cases = {"missing.png": None, "notpng.png": b"P6 4 4 255 " + bytes(48), "cut.png": blob[:len(blob) // 2],
             "garbled.png": blob[:k] + bytes([blob[k] ^ 0x55]) + blob[k + 1:]}
    for name, data in cases.items():
        if data is not None:
            (tmp_path / name).write_bytes(data)
        with pytest.raises(RuntimeError):
            scene.add_png(tmp_path / name)
    assert capfd.readouterr().err.count("cannot load PNG file") == len(cases)
    k = blob.index(b"tEXt") + 6                                             # damage in an ancillary chunk is not an error (libpng warns)
    (tmp_path / "comment.png").write_bytes(blob[:k] + b"?" + blob[k + 1:])
    assert scene.add_png(tmp_path / "comment.png") == 2
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    assert scene.view.num_textures == 1                                     # only the dummy of "none.xyz"


# ---- the JPEG loader ------------------------------------------------------------------------------------------
def photo(h, w, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(x / 7.0) * np.cos(y / 5.0), 128 + 90 * np.cos(x / 3.0 + y / 11.0), 60 + x * 2 % 200], 2)
    return np.clip(img + rng.normal(0, 20, img.shape), 0, 255).astype(np.uint8)


def jpg_expected(path):
    """What load_jpg leaves (image.cpp:186-238) from libjpeg's pixels, which PIL (libjpeg-turbo) provides: rows bottom-up,
    gamma on the colour bytes, alpha 0; a grey file fills only the first byte of each pixel."""
    lut = gamma_lut().astype(np.uint32)
    im = Image.open(path)
    px = np.array(im)[::-1].astype(np.uint32)
    return lut[px] if im.mode == "L" else lut[px[..., 0]] | lut[px[..., 1]] << 8 | lut[px[..., 2]] << 16


@pytest.mark.parametrize("h,w", [(64, 64), (37, 53), (8, 8), (1, 1), (17, 16), (100, 3), (3, 100), (33, 2), (250, 130)])
def test_jpg_loader_matches_libjpeg(tmp_path, h, w):
    """Pixel for pixel: integer IDCT, fancy chroma upsampling (replication when the chroma plane is at most two samples
    wide), fixed-point colour conversion; 4:4:4 / 4:2:2 / 4:2:0, custom Huffman tables, restart intervals, sequential and
    progressive (spectral selection + successive approximation, one AC scan per component)."""
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    for sub in (0, 1, 2):
        for k, kw in enumerate((dict(quality=50), dict(quality=95, optimize=True), dict(quality=75, restart_marker_blocks=3),
                                dict(quality=30, restart_marker_rows=1), dict(quality=60, progressive=True),
                                dict(quality=92, progressive=True, restart_marker_blocks=2), dict(quality=15, progressive=True))):
            path = tmp_path / f"s{sub}_{k}.jpg"
            Image.fromarray(photo(h, w, sub * 4 + k)).save(path, subsampling=sub, **kw)
            if k in (2, 3, 5):
                assert b"\xff\xdd" in path.read_bytes()                     # a DRI segment: the restart path is exercised
            if k >= 4:
                assert b"\xff\xc2" in path.read_bytes()                     # SOF2
            tid = scene.add_png(path)
            tex = scene.array("textures")[tid - 1]
            assert (tex["width"], tex["height"]) == (w, h)
            got = scene.array("texture_pixels")[tex["offset"]:tex["offset"] + w * h].reshape(h, w)
            assert np.array_equal(got, jpg_expected(path)), (sub, kw)


def test_jpg_grey_and_errors(tmp_path, capfd):
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    for k, kw in enumerate((dict(quality=80), dict(quality=80, progressive=True))):
        Image.fromarray(photo(40, 50)[..., 0]).save(tmp_path / "grey.jpg", **kw)
        tid = scene.add_png(tmp_path / "grey.jpg")
        tex = scene.array("textures")[tid - 1]
        got = scene.array("texture_pixels")[tex["offset"]:tex["offset"] + 2000].reshape(40, 50)
        assert np.array_equal(got, jpg_expected(tmp_path / "grey.jpg")) and (got >> 8 == 0).all()  # (grey, 0, 0, 0): image.cpp:226-227
    blob = (tmp_path / "grey.jpg").read_bytes()
    (tmp_path / "prog.jpg").write_bytes(blob.replace(b"\xff\xc2", b"\xff\xc9", 1))      # SOF9: arithmetic coding, no decoder here
    (tmp_path / "cut.jpg").write_bytes(blob[:blob.index(b"\xff\xda")])              # no scan at all (a scan cut short decodes, as in libjpeg)
    (tmp_path / "not.jpg").write_bytes(b"GIF89a" + bytes(100))
    for name in ("prog.jpg", "cut.jpg", "not.jpg", "missing.jpg"):
        with pytest.raises(RuntimeError):
            scene.add_png(tmp_path / name)
    err = capfd.readouterr().err
    assert "this JPEG coding process is not supported" in err and err.count("cannot load JPG file") == 4
    # through an OBJ: a file this decoder cannot read falls back to the constant with a warning, a missing one fails the load
    Image.fromarray(photo(16, 16)).save(tmp_path / "floor.jpg", quality=90)
    scene = R.Scene.load_obj(write_scene(tmp_path, "floor.jpg", "prog.jpg"))
    mats = scene.array("materials")
    assert mats["map_kd"].tolist()[0] == 1 and mats["map_ks"].tolist()[1] == 0
    assert np.array_equal(scene.array("texture_pixels").reshape(16, 16), jpg_expected(tmp_path / "floor.jpg"))
    assert "this JPEG coding process is not supported; the material's constant colour is used instead" in capfd.readouterr().err
    with pytest.raises(RuntimeError):
        R.Scene.load_obj(write_scene(tmp_path, "floor.jpg", "nowhere.jpeg"))


# ---- the TGA loader --------------------------------------------------------------------------------------------
def test_tga_loader(tmp_path, capfd):
    """Types 1 / 2 / 3 and their run-length forms, both row orders, against PIL's reading of the same files; 5-5-5 pixels by
    hand.  (The reference names a load_tga, converter.cpp:759-762, but no device defines it.)"""
    import struct
    scene = R.Scene.load_obj(write_scene(tmp_path, "none.xyz", "none.xyz"))
    rng = np.random.default_rng(0)
    h, w = 19, 27
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    img[:, :9] = img[:, :1]                                                # runs, so that the RLE packets are of both kinds
    for mode in ("RGBA", "RGB", "L", "P"):
        for kw in (dict(), dict(compression="tga_rle"), dict(orientation=-1), dict(compression="tga_rle", orientation=-1)):
            im = Image.fromarray(img) if mode == "RGBA" else Image.fromarray(img[..., 0]) if mode == "L" else \
                Image.fromarray(img[..., :3]).convert(mode)
            path = tmp_path / "t.tga"
            im.save(path, **kw)
            tid = scene.add_png(path)
            tex = scene.array("textures")[tid - 1]
            got = scene.array("texture_pixels")[tex["offset"]:tex["offset"] + w * h].reshape(h, w)
            assert np.array_equal(got, expected_pixels(np.array(Image.open(path).convert("RGBA")))), (mode, kw)
    # 16-bit true colour, bottom row first: 5 bits per channel widened by bit replication
    vals = np.array([[0x7C00, 0x03E0], [0x001F, 0x4210]], np.uint16)        # red, green / blue, mid grey (16, 16, 16)
    (tmp_path / "w.tga").write_bytes(struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, 2, 2, 16, 0) + vals.tobytes())
    tid = scene.add_png(tmp_path / "w.tga")
    tex = scene.array("textures")[tid - 1]
    got = scene.array("texture_pixels")[tex["offset"]:tex["offset"] + 4].reshape(2, 2)
    lut = gamma_lut().astype(np.uint32)
    g = lut[132]
    assert got.tolist() == [[0xFF000000 | 255, 0xFF000000 | 255 << 8], [0xFF000000 | 255 << 16, 0xFF000000 | g | g << 8 | g << 16]]
    blob = (tmp_path / "t.tga").read_bytes()
    (tmp_path / "cut.tga").write_bytes(blob[:40])
    (tmp_path / "odd.tga").write_bytes(blob[:2] + bytes([32]) + blob[3:])
    for name in ("cut.tga", "odd.tga", "missing.tga"):
        with pytest.raises(RuntimeError):
            scene.add_png(tmp_path / name)
    assert capfd.readouterr().err.count("cannot load TGA file") == 3


# ---- binding images to materials --------------------------------------------------------------------------------
def test_obj_binds_images_to_materials(tmp_path, capfd):
    Image.fromarray(checker()).save(tmp_path / "floor.png")
    (tmp_path / "sub").mkdir()
    Image.fromarray(checker(8, 1, 5)).save(tmp_path / "sub" / "wall.png")
    scene = R.Scene.load_obj(write_scene(tmp_path, "floor.png", "sub\\wall.png"))         # '\' -> '/', converter.cpp:84-91
    mats = scene.array("materials")
    # complex materials first (converter.cpp:460-465, 545): textured and emissive ones are all "complex", file order kept
    assert mats["bsdf"].tolist() == [R.BSDF_DIFFUSE, R.BSDF_MIX, R.BSDF_DIFFUSE]
    assert mats["map_kd"].tolist() == [1, 0, 0] and mats["map_ks"].tolist() == [0, 2, 0]
    tex = scene.array("textures")
    assert tex["width"].tolist() == [16, 8] and tex["offset"].tolist() == [0, 256]
    assert scene.view.num_texture_pixels == 256 + 64
    assert np.array_equal(scene.array("texture_pixels")[:256].reshape(16, 16), expected_pixels(checker()))
    tc = scene.array("texcoords")[scene.array("indices")[0, :3], :2]
    assert tc.tolist() == [[0, 0], [3, 0], [3, 3]]

    # one file used twice is one image; an unknown extension is the reference's black 1x1 dummy (converter.cpp:746, 765-767);
    # a format without a decoder here falls back to the constant with a warning
    scene = R.Scene.load_obj(write_scene(tmp_path, "floor.png", "floor.png"))
    assert scene.array("materials")["map_kd"].tolist()[0] == 1 and scene.array("materials")["map_ks"].tolist()[1] == 1
    assert scene.view.num_textures == 1
    scene = R.Scene.load_obj(write_scene(tmp_path, "floor.bmp", "wall.tiff"))
    err = capfd.readouterr().err
    assert "no decoder for 'wall.tiff'" in err
    assert scene.array("materials")["map_kd"].tolist()[0] == 1 and scene.array("materials")["map_ks"].tolist()[1] == 0
    assert scene.array("texture_pixels").tolist() == [0xFF000000]
    # a PNG that cannot be read fails the load, as the reference's load_png does at start-up (interface.cpp:476-477)
    with pytest.raises(RuntimeError):
        R.Scene.load_obj(write_scene(tmp_path, "floor.png", "nowhere.png"))
    assert "cannot load PNG file" in capfd.readouterr().err


# ---- the oracle's texturing ------------------------------------------------------------------------------------
def test_oracle_constant_texture_equals_constant_colour(tmp_path):
    """A one-colour image is the constant colour lut[p] / 255 up to the rounding of lerp(a, a, k): the film of the textured
    scene and of the scene with that Kd / Ks written into the MTL agree to 1e-4 of the mean, and nearly all pixels to 1e-5."""
    p = np.array([200, 100, 50], np.uint8)
    const = np.zeros((4, 4, 4), np.uint8)
    const[..., :3], const[..., 3] = p, 255
    Image.fromarray(const).save(tmp_path / "floor.png")
    Image.fromarray(const).save(tmp_path / "wall.png")
    k = gamma_lut()[p].astype(np.float32) * np.float32(1.0 / 255.0)
    textured = R.Scene.load_obj(write_scene(tmp_path))
    (tmp_path / "flat").mkdir()
    (tmp_path / "flat" / "tex.obj").write_text(OBJ)
    kstr = " ".join(repr(float(x)) for x in k)
    flat_mtl = MTL.format(floor="", wall="", wall_kd="0.4 0.3 0.2").replace("map_Kd \n", "").replace("map_Ks \n", "")
    flat_mtl = flat_mtl.replace("Kd 0.5 0.5 0.5", "Kd " + kstr).replace("Ks 0.3 0.3 0.3", "Ks " + kstr)
    (tmp_path / "flat" / "tex.mtl").write_text(flat_mtl)
    flat = R.Scene.load_obj(tmp_path / "flat" / "tex.obj")
    assert flat.view.num_textures == 0 and textured.view.num_textures == 2
    assert np.array_equal(flat.array("materials")["kd"][1], k)             # simple materials follow the emitter: lamp, floor, wall
    W, H = 48, 32
    a, _ = oracle.render(textured.view, camera(W, H), W, H, 4, 6, 0)
    b, _ = oracle.render(flat.view, camera(W, H), W, H, 4, 6, 0)
    assert b.mean() > 0.05
    rel = np.abs(a - b) / (np.abs(b) + 1e-3)
    assert abs(a.mean() - b.mean()) / b.mean() < 1e-4 and (rel < 1e-5).mean() > 0.99, (a.mean(), b.mean(), (rel < 1e-5).mean())


def test_oracle_sees_the_checker(tmp_path):
    """Looking straight down at the floor (uv 0..3 over the quad, 16 texels of 2x2-texel cells per repeat): the primary-hit
    albedo changes the film, so bright and dark cells show up as a strong variation along a floor row; with the texture
    replaced by its average colour that variation is gone."""
    img = checker(16, 8, 1)
    img[..., :3] = np.where(img[..., :1] > 0, img[..., :3] // 2 + 120, 0)
    dark = (np.add.outer(np.arange(16) // 8, np.arange(16) // 8) % 2) == 0
    img[dark, :3] = 8
    Image.fromarray(img).save(tmp_path / "floor.png")
    Image.fromarray(img).save(tmp_path / "wall.png")
    scene = R.Scene.load_obj(write_scene(tmp_path))
    W, H = 64, 48
    film = np.zeros((H, W, 3), np.float32)
    for it in range(4):
        film, _ = oracle.render(scene.view, camera(W, H), W, H, 8, 4, it, film)
    floor_row = film[H - 6].sum(axis=1)
    smooth = np.convolve(floor_row, np.ones(3) / 3, "valid")
    assert smooth.max() > 3.0 * smooth.min(), (smooth.max(), smooth.min())


# ---- emission textures (MTL map_Ke, converter.cpp:794-801) -----------------------------------------------------
LAMP_OBJ = OBJ.replace("f 7 8 9 10", "f 7/1 8/2 9/3 10/4")        # the lamp quad gets texture coordinates 0..3 as well


def write_lamp_scene(tmp_path, ke_line, map_ke):
    (tmp_path / "tex.obj").write_text(LAMP_OBJ)
    mtl = MTL.format(floor="", wall="", wall_kd="0.4 0.3 0.2").replace("map_Kd \n", "").replace("map_Ks \n", "")
    mtl = mtl.replace("Ke 12 12 10\n", ke_line + (f"map_Ke {map_ke}\n" if map_ke else ""))
    (tmp_path / "tex.mtl").write_text(mtl)
    return tmp_path / "tex.obj"


def test_map_ke_binds_to_the_material_and_its_lights(tmp_path):
    Image.fromarray(checker(8, 2, 4)).save(tmp_path / "glow.png")
    scene = R.Scene.load_obj(write_lamp_scene(tmp_path, "Ke 0 0 0\n", "glow.png"))
    mats, lights = scene.array("materials"), scene.array("lights")
    assert scene.view.num_textures == 1 and scene.view.num_lights == 2      # an emitter through its texture alone
    assert mats["is_emissive"].tolist() == [1, 0, 0] and mats["map_ke"].tolist() == [1, 0, 0]
    idx = scene.array("indices")
    for l in lights:
        assert l["map_ke"] == 1 and idx[l["prim"], 3] == 0                   # the light knows its triangle
        assert np.array_equal(scene.array("vertices")[idx[l["prim"], 0], :3], l["v0"])


def test_oracle_constant_emission_texture_equals_constant_ke(tmp_path):
    """A one-colour map_Ke is the constant Ke = lut[p] / 255: same film up to the rounding of the bilinear lerp."""
    p = np.array([250, 180, 90], np.uint8)
    const = np.zeros((4, 4, 4), np.uint8)
    const[..., :3], const[..., 3] = p, 255
    Image.fromarray(const).save(tmp_path / "glow.png")
    k = gamma_lut()[p].astype(np.float32) * np.float32(1.0 / 255.0)
    textured = R.Scene.load_obj(write_lamp_scene(tmp_path, "Ke 0 0 0\n", "glow.png"))
    (tmp_path / "flat").mkdir()
    flat = R.Scene.load_obj(write_lamp_scene(tmp_path / "flat", "Ke " + " ".join(repr(float(x)) for x in k) + "\n", None))
    W, H = 48, 32
    a, _ = oracle.render(textured.view, camera(W, H), W, H, 4, 6, 0)
    b, _ = oracle.render(flat.view, camera(W, H), W, H, 4, 6, 0)
    assert b.mean() > 0.005
    rel = np.abs(a - b) / (np.abs(b) + 1e-3)
    assert abs(a.mean() - b.mean()) / b.mean() < 1e-4 and (rel < 1e-5).mean() > 0.99, (a.mean(), b.mean(), (rel < 1e-5).mean())


def test_oracle_emission_texture_shapes_the_light(tmp_path):
    """Half of the emission image black: the scene receives about half the light of the all-bright image."""
    bright = np.full((8, 8, 4), 255, np.uint8)
    half = bright.copy()
    half[:, :4, :3] = 0
    films = {}
    for name, img in (("bright", bright), ("half", half)):
        d = tmp_path / name
        d.mkdir()
        Image.fromarray(img).save(d / "glow.png")
        scene = R.Scene.load_obj(write_lamp_scene(d, "Ke 0 0 0\n", "glow.png"))
        film = np.zeros((32, 48, 3), np.float32)
        for it in range(4):
            film, _ = oracle.render(scene.view, camera(48, 32), 48, 32, 8, 4, it, film)
        films[name] = film
    ratio = films["half"][20:].mean() / films["bright"][20:].mean()          # the floor rows: lit by next-event estimation
    assert 0.35 < ratio < 0.65, ratio


@pytest.mark.gpu
def test_emission_texture_film_matches_oracle(tmp_path):
    Image.fromarray(checker(8, 2, 4)).save(tmp_path / "glow.png")
    scene = R.Scene.load_obj(write_lamp_scene(tmp_path, "Ke 0 0 0\n", "glow.png"))
    W, H, spp, depth = 120, 90, 4, 6
    cam = camera(W, H)
    r = R.Renderer(scene, 0, W, H, spp, depth)
    want = np.zeros((H, W, 3), np.float32)
    for it in range(2):
        r.render(cam, it)
        want, _ = oracle.render(scene.view, cam, W, H, spp, depth, it, want)
    got = r.film().copy()
    r.free()
    e = np.abs(got - want) / (np.abs(want) + 1e-3)
    assert want.mean() > 0.002
    assert np.median(e) < 1e-5 and (e < 1e-3).mean() > 0.995, (np.median(e), (e < 1e-3).mean())


# ---- CUDA ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("size,spp,depth", [((160, 120), 4, 8), ((67, 45), 3, 32)])
def test_textured_film_matches_oracle(tmp_path, size, spp, depth):
    """Same tolerance as tests/test_gpu_render.py::test_film_matches_oracle (sinf / cosf last bits, atomics order)."""
    Image.fromarray(checker(16, 2, 3)).save(tmp_path / "floor.png")
    Image.fromarray(checker(8, 1, 5)).save(tmp_path / "wall.png")
    scene = R.Scene.load_obj(write_scene(tmp_path))
    W, H = size
    cam = camera(W, H)
    r = R.Renderer(scene, 0, W, H, spp, depth)
    want = np.zeros((H, W, 3), np.float32)
    for it in range(2):
        r.render(cam, it)
        want, st = oracle.render(scene.view, cam, W, H, spp, depth, it, want)
    got = r.film().copy()
    stats = r.stats()
    r.free()
    e = np.abs(got - want) / (np.abs(want) + 1e-3)
    assert np.median(e) < 1e-5 and (e < 1e-3).mean() > 0.995, (np.median(e), (e < 1e-3).mean())
    assert abs(got.mean() - want.mean()) / want.mean() < 2e-3
    assert abs(stats["primary_rays"] - st.primary_rays) <= 0.001 * st.primary_rays + 4
    # and the texture matters: the same scene without its images renders something else
    flat = R.Scene.load_obj(write_scene(tmp_path, "floor.tiff", "wall.tiff"))
    r = R.Renderer(flat, 0, W, H, spp, depth)
    for it in range(2):
        r.render(cam, it)
    other = r.film().copy()
    r.free()
    assert np.abs(other - got).mean() > 0.02 * got.mean()


@pytest.mark.gpu
def test_texture_added_to_a_bvh_scene(tmp_path):
    """rodent_b200_scene_add_texture on a scene that was not built from an OBJ; a texture id out of range is refused."""
    from rodent_b200 import formats, testdata, workloads
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
    mats, mop = workloads.sponza_materials(), workloads.sponza_material_of_prim(tris)
    W, H = 96, 64
    cam = workloads.camera("sponza", W, H)
    mats["map_kd"][mats["bsdf"] == R.BSDF_DIFFUSE] = 1
    scene = R.Scene.from_bvh8(nodes, tris, mats, mop)
    with pytest.raises(RuntimeError):
        R.Renderer(scene, 0, W, H, 1, 4)                                    # texture 1 does not exist yet
    tid = scene.add_texture(expected_pixels(checker(4, 1, 9)))
    assert tid == 1
    r = R.Renderer(scene, 0, W, H, 2, 4)
    r.render(cam, 0)
    want, _ = oracle.render(scene.view, cam, W, H, 2, 4, 0)
    got = r.film().copy()
    r.free()
    e = np.abs(got - want) / (np.abs(want) + 1e-3)
    assert (e < 1e-3).mean() > 0.99
