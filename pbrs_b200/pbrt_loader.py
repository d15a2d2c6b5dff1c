"""pbrt-v3-subset scene files -> SceneDesc: the input side of the boundary.

Mirrors what the reference does between a scene file and its `Scene`:
  lexer    scene_parser/src/token.rs (token set, no exponent floats :112-114, `#` comments) and
           scene_parser/src/lexer.rs:41-56 (`Include` splices the tokens of another file)
  parser   scene_parser/src/parser.rs (scene-wide options, WorldBegin ... WorldEnd, attribute /
           transform blocks, parameter lists; a one-element list is a Number, :230-239)
  loader   scene/src/loader.rs: camera (:91-135), world items (:164-305), shapes sphere /
           trianglemesh (:307-389), area-light shapes (:396-434, sphere only: the PLY loader is
           truncated upstream), lights distant / point / infinite (:257-284, :436-481), materials
           glass / mirror / matte / metal / plastic / uber / substrate (:483-714), textures
           imagemap (:716-731), constant colours rgb / color (:758-766).
Transforms are composed in FP32 exactly as the reference does (AffineTransform = forward and
inverse Mat4 multiplied separately, geometry/src/transform.rs:185-194; `Rotate` uses the negated
angle, loader.rs:792-798; Mat4 x Mat4 column by column, math/src/hcm.rs:546-556).

Unsupported directives raise `PbrtError` where the reference panics / hits `unimplemented!()`
(plymesh: scene/src/plyloader.rs is cut off upstream; its tail is restated here, see load_ply.
loopsubdiv, fourier, spectrum / xyz / blackbody colours, ObjectBegin/ObjectInstance: out of scope,
DESIGN.md section 8).
"""
import os
import re

import numpy as np

from . import _capi as K
from .scene import SceneDesc

F32 = np.float32


class PbrtError(ValueError):
    pass


KEYWORDS = {
    "Include", "LookAt", "Camera", "Integrator", "Accelerator", "Sampler", "Film", "PixelFilter", "Filter", "WorldBegin", "WorldEnd",
    "AttributeBegin", "AttributeEnd", "TransformBegin", "TransformEnd", "LightSource", "AreaLightSource", "Material", "Shape", "Texture",
    "Identity", "Translate", "Scale", "Rotate", "CoordinateSystem", "CoordSysTransform", "Transform", "ConcatTransform", "ReverseOrientation",
    "MediumInterface", "NamedMedium", "MakeNamedMedium", "NamedMaterial", "MakeNamedMaterial", "ObjectBegin", "ObjectEnd", "ObjectInstance",
}
# scene_parser/src/token.rs: whitespace and comments are skipped; floats have NO exponent form
_TOKEN = re.compile(r"""[ \t\n\f\r]+ | \#[^\n]*\n? | (?P<lb>\[) | (?P<rb>\]) | "(?P<str>[^"\n]+)" |
                        (?P<num>[-+]?(?:\d+(?:\.\d*)?|\.\d+)) | (?P<kw>[A-Za-z]+)""", re.X)
TRANSFORM_START = {"Identity", "Translate", "Scale", "Rotate", "LookAt", "Transform", "ConcatTransform", "CoordSysTransform", "CoordinateSystem"}
SCENE_OPTION_START = {"Camera", "Sampler", "Film", "Filter", "Integrator", "Accelerator"} | TRANSFORM_START
WORLD_ITEM_START = TRANSFORM_START | {"Shape", "Material", "LightSource", "AreaLightSource", "Texture", "MakeNamedMaterial", "ObjectInstance",
                                      "AttributeBegin", "ObjectBegin", "TransformBegin", "NamedMaterial", "ReverseOrientation"}


def tokenize(text, root_dir="."):
    """-> list of ('kw', name) | ('num', float32) | ('str', s) | ('[',) | (']',)"""
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise PbrtError(f"lexer error at {text[pos:pos + 30]!r}")  # Token::Error
        pos = m.end()
        if m.group("lb"):
            out.append(("[",))
        elif m.group("rb"):
            out.append(("]",))
        elif m.group("str") is not None:
            out.append(("str", m.group("str")))
        elif m.group("num") is not None:
            out.append(("num", F32(m.group("num"))))  # str::parse::<f32>
        elif m.group("kw") is not None:
            if m.group("kw") not in KEYWORDS:
                raise PbrtError(f"unknown directive {m.group('kw')!r}")
            out.append(("kw", m.group("kw")))
    # lexer.rs:41-56: Include "file" splices that file's tokens
    res, i = [], 0
    while i < len(out):
        if out[i] == ("kw", "Include"):
            if i + 1 >= len(out) or out[i + 1][0] != "str":
                raise PbrtError("should have a file name after Include")
            path = os.path.join(root_dir, out[i + 1][1])
            res.extend(tokenize(open(path).read(), os.path.dirname(path)))
            i += 2
        else:
            res.append(out[i])
            i += 1
    return res


# ---- FP32 matrices, math/src/hcm.rs (column vectors of a column-major Mat4) ----
def _libm_f32(name):
    """f32::sin / f32::cos are the platform libm's sinf / cosf (what the C++ loader calls too); numpy's
    float32 kernels are a different implementation and differ from them in the last bit now and then."""
    import ctypes
    import ctypes.util
    try:
        fn = getattr(ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6"), name)
        fn.restype = ctypes.c_float
        fn.argtypes = [ctypes.c_float]
        return lambda x: F32(fn(float(F32(x))))
    except (OSError, AttributeError):
        import math
        ref = {"sinf": math.sin, "cosf": math.cos}[name]
        return lambda x: F32(ref(float(F32(x))))   # correctly rounded double, rounded once more


_sinf, _cosf = _libm_f32("sinf"), _libm_f32("cosf")


def _ident():
    return np.eye(4, dtype=F32)


def _mat_vec(m, v):  # hcm.rs:539-544: ((c0*v0 + c1*v1) + c2*v2) + c3*v3
    return ((m[:, 0] * v[0] + m[:, 1] * v[1]) + m[:, 2] * v[2]) + m[:, 3] * v[3]


def _mat_mul(a, b):  # hcm.rs:546-556: mat.cols[c] = ZERO + a * b.cols[c]
    out = np.zeros((4, 4), F32)
    for c in range(4):
        out[:, c] = np.zeros(4, F32) + _mat_vec(a, b[:, c])
    return out


class Affine:
    """geometry/src/transform.rs:16-19: forward and inverse carried side by side."""

    def __init__(self, fwd=None, inv=None):
        self.fwd = _ident() if fwd is None else fwd
        self.inv = _ident() if inv is None else inv

    def __mul__(self, rhs):  # transform.rs:185-194
        return Affine(_mat_mul(self.fwd, rhs.fwd), _mat_mul(rhs.inv, self.inv))

    def is_identity(self):
        return np.array_equal(self.fwd, _ident()) and np.array_equal(self.inv, _ident())

    @staticmethod
    def translater(t):  # transform.rs:140-145
        f, i = _ident(), _ident()
        f[:3, 3] = t
        i[:3, 3] = -np.asarray(t, F32)
        return Affine(f, i)

    @staticmethod
    def scaler(s):  # transform.rs:159-166
        f, i = _ident(), _ident()
        for k in range(3):
            f[k, k] = s[k]
            i[k, k] = F32(1.0) / F32(s[k])
        return Affine(f, i)

    @staticmethod
    def rotater(axis, angle_rad):  # transform.rs:146-152 over hcm.rs:508-520
        axis = np.asarray(axis, F32)
        sin_t, cos_t = _sinf(angle_rad), _cosf(angle_rad)
        f = _ident()
        dot = lambda a, b: F32(F32(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])
        ahat = axis * (F32(1.0) / F32(np.sqrt(dot(axis, axis))))
        for i in range(3):
            base = np.zeros(3, F32)
            base[i] = 1.0
            vc = dot(base, axis) * axis / dot(axis, axis)
            v1 = base - vc
            v2 = np.array([v1[1] * ahat[2] - v1[2] * ahat[1], v1[2] * ahat[0] - v1[0] * ahat[2], v1[0] * ahat[1] - v1[1] * ahat[0]], F32)
            f[:3, i] = vc + v1 * cos_t + v2 * sin_t
        return Affine(f, f.T.copy())

    def apply_point(self, p):
        return _mat_vec(self.fwd, np.array([p[0], p[1], p[2], 1.0], F32))[:3]

    def apply_vec(self, v):
        return _mat_vec(self.fwd, np.array([v[0], v[1], v[2], 0.0], F32))[:3]


# ---------------------------------------------------------------------------------------------
# PLY meshes: scene/src/plyloader.rs:69-256 (binary only; float vertex properties; polygons are
# fanned) and geometry/src/lib.rs:16-32 (compute_normals).  The file is cut off upstream right
# after the normals are computed; the tail restated here is what the signature and
# the call sites (loader.rs:314-331,408-419) require: Vertex{pos, normal, uv} with uv defaulting
# to (0, 0), then TriangleMeshRaw{vertices, index_triples}.
# ---------------------------------------------------------------------------------------------
_PLY_SIZES = {"uchar": 1, "uint8": 1, "short": 2, "int": 4, "uint": 4}  # get_type_size, :14-21


def _cross32(a, b):  # hcm.rs:89-98, per row, in f32
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(F32)


def compute_normals(P, idx):
    """geometry/src/lib.rs:16-32: face normals (p1-p0)x(p2-p0) accumulated per vertex in triangle
    order (f32), then Vec3::hat.  A vertex no triangle touches makes hat() panic upstream."""
    P = np.asarray(P, F32)
    n = _cross32(P[idx[:, 1]] - P[idx[:, 0]], P[idx[:, 2]] - P[idx[:, 0]])
    acc = np.zeros_like(P)
    np.add.at(acc, idx.reshape(-1), np.repeat(n, 3, axis=0))  # unbuffered, in (i, j, k) order per triangle
    n2 = ((acc[:, 0] * acc[:, 0] + acc[:, 1] * acc[:, 1]) + acc[:, 2] * acc[:, 2]).astype(F32)
    if not (np.isfinite(n2) & (n2 != 0)).all():
        raise PbrtError("compute_normals: a vertex has no (finite, non-zero) normal; Vec3::hat panics upstream")
    inv = (F32(1.0) / np.sqrt(n2)).astype(F32)
    return (acc * inv[:, None]).astype(F32)


def load_ply(path):
    """-> (P[n,3], N[n,3], UV[n,2], idx[m,3]) as the reference's TriangleMeshRaw would hold them."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def line():
        nonlocal pos
        end = data.find(b"\n", pos)
        if end < 0:
            raise PbrtError("ply: unexpected end of header")
        out = data[pos:end + 1].decode("ascii", "replace")
        pos = end + 1
        return out

    if line().strip() != "ply":
        raise PbrtError("ply: Header isn't ply")
    words = line().split(" ")
    if words[0] != "format" or len(words) < 2:
        raise PbrtError("ply: Format line is bad")
    fmt = words[1].strip()
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise PbrtError(f"ply: Unrecognized format string: {fmt}")
    props, nv, nf, len_size, el_size = [], None, None, None, None
    while True:
        ln = line().strip()
        if ln == "end_header":
            break
        if ln.startswith("comment"):
            continue
        w = ln.split(" ")
        if len(w) < 3:
            raise PbrtError(f"ply: Can't handle the line {ln}")
        if w[:2] == ["element", "vertex"] and len(w) == 3:
            nv = int(w[2]) if (w[2].isdigit() and len(w[2]) <= 10 and int(w[2]) < 2 ** 31) else None  # same limit as the C++ loader
        elif w[:2] == ["element", "face"] and len(w) == 3:
            nf = int(w[2]) if (w[2].isdigit() and len(w[2]) <= 10 and int(w[2]) < 2 ** 31) else None
        elif w[:2] == ["property", "float"] and len(w) == 3:
            props.append(w[2])
        elif w[:2] == ["property", "list"] and len(w) == 5 and w[4] == "vertex_indices":
            len_size, el_size = _PLY_SIZES.get(w[2]), _PLY_SIZES.get(w[3])
        elif w[0] == "property":
            # upstream only warns and then mis-reads the vertex block (its size assumes floats only)
            raise PbrtError(f"ply: unsupported property line '{ln}' (only float vertex properties and a vertex_indices list)")
        # other element kinds: "Unprocessed header line" upstream
    if None in (nv, nf, len_size, el_size):
        raise PbrtError("ply: header lacks vertex / face counts or the vertex_indices list")
    if fmt == "ascii":
        raise PbrtError("ply: ascii payloads are not supported (bytes_to_f32 panics upstream)")
    e = "<" if fmt == "binary_little_endian" else ">"
    stride = len(props)
    nbytes = nv * stride * 4
    if pos + nbytes > len(data):
        raise PbrtError("ply: vertex block is truncated")
    vb = np.frombuffer(data, dtype=e + "f4", count=nv * stride, offset=pos).astype(F32).reshape(nv, stride)
    pos += nbytes
    tris = []
    for _ in range(nf):
        if pos + len_size > len(data):
            raise PbrtError("ply: face block is truncated")
        n = int.from_bytes(data[pos:pos + len_size], "little" if e == "<" else "big")
        pos += len_size
        if pos + n * el_size > len(data):
            raise PbrtError("ply: face block is truncated")
        if n == 0:
            raise PbrtError("ply: empty face (list_length - 1 underflows upstream)")
        face = np.frombuffer(data, dtype=e + {1: "u1", 2: "u2", 4: "u4"}[el_size], count=n, offset=pos).astype(np.int64)
        pos += n * el_size
        if n == 3:
            tris.append((face[0], face[1], face[2]))
        else:  # fan, :184-190
            for i in range(1, n - 1):
                tris.append((face[0], face[i], face[i + 1]))
    idx = np.array(tris, np.int64).reshape(-1, 3)
    if idx.size and (idx.min() < 0 or idx.max() >= nv):
        raise PbrtError("ply: vertex index out of range")
    off = {name: i for i, name in enumerate(props)}  # the last occurrence wins, as in the match loop (:211-223)
    if not all(k in off for k in ("x", "y", "z")):
        raise PbrtError("ply: position xyz: some missing")
    P = vb[:, [off["x"], off["y"], off["z"]]].copy()
    idx = idx.astype(np.uint32)
    if all(k in off for k in ("nx", "ny", "nz")):
        N = vb[:, [off["nx"], off["ny"], off["nz"]]].copy()
    else:
        N = compute_normals(P, idx)
    UV = vb[:, [off["u"], off["v"]]].copy() if ("u" in off and "v" in off) else np.zeros((nv, 2), F32)
    return P, N, UV, idx


def write_ply(path, P, idx, N=None, UV=None, big_endian=False, index_type="int", polygons=None):
    """Binary PLY writer for fixtures (the subset load_ply reads).  `polygons`: optional list of
    index lists written instead of `idx` (to exercise the fan triangulation)."""
    P = np.asarray(P, F32)
    cols, names = [P], ["x", "y", "z"]
    if N is not None:
        cols.append(np.asarray(N, F32)); names += ["nx", "ny", "nz"]
    if UV is not None:
        cols.append(np.asarray(UV, F32)); names += ["u", "v"]
    faces = [list(map(int, t)) for t in (polygons if polygons is not None else np.asarray(idx).reshape(-1, 3))]
    e = ">" if big_endian else "<"
    it = {"uchar": "u1", "short": "u2", "int": "u4", "uint": "u4"}[index_type]
    with open(path, "wb") as f:
        hdr = ["ply", f"format binary_{'big' if big_endian else 'little'}_endian 1.0", "comment written by pbrs_b200",
               f"element vertex {P.shape[0]}"] + [f"property float {n}" for n in names] + \
              [f"element face {len(faces)}", f"property list uchar {index_type} vertex_indices", "end_header"]
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        f.write(np.concatenate(cols, axis=1).astype(e + "f4").tobytes())
        for face in faces:
            f.write(bytes([len(face)]))
            f.write(np.asarray(face, e + it).tobytes())


def _to_radians(deg):  # f32::to_radians
    return F32(deg) * (F32(np.pi) / F32(180.0))


class _Params(dict):
    """scene_parser/src/ast.rs:14-70 ParameterSet."""

    def extract(self, key):
        return self.pop(key, None)

    def extract_substr(self, pattern):  # :58-69: first key one of whose space-separated parts == pattern
        for k in list(self):
            if pattern in k.split(" "):
                return k, self.pop(k)
        return None

    def lookup_f32(self, key):
        v = self.get(key)
        return v if isinstance(v, (F32, float)) else None


class Parser:
    """scene_parser/src/parser.rs, producing nested tuples."""

    def __init__(self, tokens):
        self.t = tokens + [("kw", "$")]
        self.i = 0

    @property
    def peek(self):
        return self.t[self.i]

    def next(self):
        self.i += 1

    def kw(self):
        return self.peek[1] if self.peek[0] == "kw" else None

    def expect_kw(self, name):
        if self.kw() != name:
            raise PbrtError(f"expected {name}, found {self.peek}")
        self.next()

    def quoted(self):
        if self.peek[0] != "str":
            raise PbrtError(f"expected a quoted string, found {self.peek}")
        s = self.peek[1]
        self.next()
        return s

    def numbers(self):
        out = []
        while self.peek[0] == "num":
            out.append(self.peek[1])
            self.next()
        return out

    def parameter_list(self):  # :249-257
        ps = _Params()
        while self.peek[0] == "str":
            key = self.quoted()
            if self.peek == ("[",):
                self.next()
                if self.peek[0] == "num":
                    nums = self.numbers()
                    val = nums[0] if len(nums) == 1 else nums  # :230-235 Number vs Numbers
                elif self.peek[0] == "str":
                    val = self.quoted()
                else:
                    raise PbrtError("only numbers or quoted strings allowed")
                if self.peek != ("]",):
                    raise PbrtError("expected ]")
                self.next()
            elif self.peek[0] == "str":
                val = self.quoted()
            elif self.peek[0] == "num":
                val = self.peek[1]
                self.next()
            else:
                raise PbrtError("unexpected token after key")
            ps[key] = val
        return ps

    def transform(self):  # :259-311
        k = self.kw()
        self.next()
        if k == "Identity":
            return ("Identity",)
        if k == "Translate":
            n = self.numbers()
            if len(n) != 3:
                raise PbrtError("wrong number of numbers after translation")
            return ("Translate", n)
        if k == "Scale":
            n = self.numbers()[:3]
            return ("Scale", n)
        if k == "Rotate":
            n = self.numbers()
            if len(n) != 4:
                raise PbrtError("4 numbers expected in Rotate")
            return ("Rotate", n[1:], _to_radians(n[0]))
        if k == "LookAt":
            n = self.numbers()
            if len(n) != 9:
                raise PbrtError("wrong numbers of floats in LookAt")
            return ("LookAt", n[0:3], n[3:6], n[6:9])
        raise PbrtError(f"unsupported transform directive {k} (unimplemented!() upstream)")

    def world_item(self):  # :38-158
        k = self.kw()
        if k in TRANSFORM_START:
            return ("Transform", self.transform())
        self.next()
        if k in ("Shape", "Material", "LightSource", "AreaLightSource"):
            impl = self.quoted()
            return (k, impl, self.parameter_list())
        if k == "Texture":
            name, ttype, impl = self.quoted(), self.quoted(), self.quoted()
            return ("Texture", impl, ttype, name, self.parameter_list())
        if k == "MakeNamedMaterial":
            name = self.quoted()
            return ("MakeNamedMaterial", name, self.parameter_list())
        if k == "NamedMaterial":
            return ("NamedMaterial", self.quoted())
        if k == "AttributeBegin":
            items = self.world_item_list()
            self.expect_kw("AttributeEnd")
            return ("AttributeBlock", items)
        if k == "TransformBegin":
            items = self.world_item_list()
            self.expect_kw("TransformEnd")
            return ("TransformBlock", items)
        if k == "ReverseOrientation":
            return ("ReverseOrientation",)
        raise PbrtError(f"unsupported world item {k} (object instancing is unimplemented!() upstream, loader.rs:781)")

    def world_item_list(self):
        items = []
        while self.kw() in WORLD_ITEM_START:
            items.append(self.world_item())
        return items

    def scene(self):  # :21-36
        options = []
        while self.kw() in SCENE_OPTION_START:
            k = self.kw()
            if k in TRANSFORM_START:
                options.append(("Transform", self.transform()))
            else:
                self.next()
                impl = self.quoted()
                options.append((k, impl, self.parameter_list()))
        self.expect_kw("WorldBegin")
        items = self.world_item_list()
        self.expect_kw("WorldEnd")
        return options, items


COPPER_ETA = (0.19547, 0.925682, 1.102186)  # preset::copper_fresnel().0, scene/src/preset.rs:478-483


class Loader:
    """scene/src/loader.rs SceneLoader, emitting SceneDesc constructor calls."""

    def __init__(self, root_dir="."):
        self.root = root_dir
        self.sd = SceneDesc()
        self.ctm = [Affine()]
        self.mtl = None          # current material id
        self.area_l = None       # current area-light luminance
        self.named_tex = {}
        self.named_mtl = {}
        self.instances = []      # (shape id, material id, Affine)

    # -- colours / textures --
    @staticmethod
    def constant_color(spectrum_type, nums):  # :758-766
        if spectrum_type in ("rgb", "color"):
            return (float(nums[0]), float(nums[1]), float(nums[2]))
        if spectrum_type == "xyz":  # Color::from_xyz, radiometry/src/color.rs:30-36 (f32, left to right)
            x, y, z = F32(nums[0]), F32(nums[1]), F32(nums[2])
            return (float(F32(3.240479) * x - F32(1.537150) * y - F32(0.498535) * z),
                    float(F32(-0.969256) * x + F32(1.875991) * y + F32(0.041556) * z),
                    float(F32(0.055648) * x - F32(0.204043) * y + F32(1.057311) * z))
        raise PbrtError(f"colour type {spectrum_type!r} is out of scope (blackbody / spectrum: load-time spectra over the CIE tables)")

    def color_arg(self, ps, name, default):
        hit = ps.extract_substr(name)
        if hit is None:
            return default
        key, val = hit
        if isinstance(val, list):
            return self.constant_color(key.split(" ")[0], val)
        if isinstance(val, str):
            raise PbrtError(f"textured {name} unsupported here (unimplemented!() upstream)")
        return (float(val),) * 3

    def solid_or_image_tex(self, key, val):  # :737-752
        if isinstance(val, list):
            return self.sd.add_texture_solid(self.constant_color(key.split(" ")[0], val))
        if isinstance(val, str):
            if val not in self.named_tex:
                raise PbrtError(f"unknown texture {val!r}")
            return self.named_tex[val]
        return self.sd.add_texture_solid((float(val),) * 3)

    def tex_arg(self, ps, name, default_gray):
        hit = ps.extract_substr(name)
        if hit is None:
            return None if default_gray is None else self.sd.add_texture_solid((default_gray,) * 3)
        return self.solid_or_image_tex(*hit)

    @staticmethod
    def num_arg(ps, name, default):
        hit = ps.extract_substr(name)
        if hit is None:
            return default
        if isinstance(hit[1], (list, str)):
            raise PbrtError(f"{name} value isn't a number: {hit[1]!r}")
        return float(hit[1])

    @staticmethod
    def bool_arg(ps, name, default):
        hit = ps.extract_substr(name)
        if hit is None:
            return default
        if hit[1] not in ("true", "false"):
            raise PbrtError(f"invalid boolean string {hit[1]!r}")
        return hit[1] == "true"

    # -- materials, :483-714 --
    def material(self, impl, ps):
        sd = self.sd
        if impl == "glass":
            kr = self.color_arg(ps, "Kr", (1.0, 1.0, 1.0))
            kt = self.color_arg(ps, "Kt", (1.0, 1.0, 1.0))
            return sd.dielectric(self.num_arg(ps, "eta", 1.5), reflect=kr, transmit=kt)
        if impl == "mirror":
            return sd.mirror(self.color_arg(ps, "Kr", (0.9, 0.9, 0.9)))
        if impl == "matte":
            kd = self.tex_arg(ps, "Kd", 0.5)
            ps.extract("sigma")  # Oren-Nayar is a TODO upstream (:532-538)
            return sd.lambertian(tex=kd)
        if impl == "metal":
            rough = self.num_arg(ps, "roughness", 0.01)
            ps.extract("remaproughness")
            eta = self.color_arg(ps, "eta", COPPER_ETA)
            k = self.color_arg(ps, "k", COPPER_ETA)  # Q16: defaults to copper ETA, :560
            return sd.metal(eta, k, rough)
        if impl == "plastic":
            kd = self.color_arg(ps, "Kd", (0.25,) * 3)
            ks = self.color_arg(ps, "Ks", (0.25,) * 3)
            rough = self.num_arg(ps, "roughness", 0.1)
            return sd.plastic(kd, ks, rough, remap_roughness=self.bool_arg(ps, "remaproughness", True))
        if impl == "uber":
            kd, ks = self.tex_arg(ps, "Kd", 0.25), self.tex_arg(ps, "Ks", 0.25)
            kr, kt = self.tex_arg(ps, "Kr", None), self.tex_arg(ps, "Kt", None)
            ur, vr = self.num_arg(ps, "uroughness", 0.0), self.num_arg(ps, "vroughness", 0.0)
            r = self.num_arg(ps, "roughness", 0.0)
            eta = self.num_arg(ps, "eta", 1.5)
            opacity = 1.0  # Q16: `opacity` re-reads "eta", already extracted -> always 1 (:644)
            remap = self.bool_arg(ps, "remaproughness", True)
            ru, rv = (r, r) if ur == vr else (ur, vr)  # Roughness::Iso / ::UV, :655-659
            return sd.uber(kd, ks, -1 if kr is None else kr, -1 if kt is None else kt, ru, rv, eta, opacity, remap)
        if impl == "substrate":
            return sd.substrate(self.tex_arg(ps, "Kd", 0.5), self.tex_arg(ps, "Ks", 0.5))
        raise PbrtError(f"not recognized material: {impl}")

    # -- shapes, :307-389 --
    def shape(self, impl, ps):
        if impl == "sphere":
            r = ps.lookup_f32("float radius")
            return self.sd.add_sphere((0.0, 0.0, 0.0), 1.0 if r is None else float(r))
        if impl == "trianglemesh":
            P = ps.extract("point P")
            if not isinstance(P, list):
                raise PbrtError("missing points")
            P = np.array(P, F32).reshape(-1, 3)
            uv = ps.extract("float uv")
            if uv is None:
                uv = ps.extract("float st")
            UV = np.zeros((P.shape[0], 2), F32) if uv is None else np.array(uv, F32).reshape(-1, 2)
            idx = ps.extract("integer indices")
            if not isinstance(idx, list):
                raise PbrtError("missing indices")
            idx = np.array(idx, F32).astype(np.uint32).reshape(-1, 3)
            nrm = ps.extract_substr("normal")
            N = np.zeros_like(P) if nrm is None else np.array(nrm[1], F32).reshape(-1, 3)
            return self.sd.add_mesh(P, idx, N=N, UV=UV)
        if impl == "plymesh":  # :314-331
            P, N, UV, idx = self.ply(ps)
            return self.sd.add_mesh(P, idx, N=N, UV=UV)
        raise PbrtError(f"shape of {impl} is out of scope (loopsubdiv: pre-process)")

    def ply(self, ps):
        name = ps.get("string filename")  # ParameterSet::lookup_string
        if not isinstance(name, str):
            raise PbrtError("no ply file specified")
        return load_ply(os.path.join(self.root, name))

    @staticmethod
    def transform_of(t):  # :784-803
        if t[0] == "Identity":
            return Affine()
        if t[0] == "Translate":
            return Affine.translater(np.array(t[1], F32))
        if t[0] == "Scale":
            return Affine.scaler(np.array(t[1], F32))
        if t[0] == "Rotate":
            return Affine.rotater(t[1], -t[2])  # the negated angle, :792-798
        raise PbrtError("unsupported lookat in modeling step")

    def light(self, impl, ps):  # :257-284, :436-481
        if impl == "infinite":
            hit = ps.extract_substr("L")
            mult = None if hit is None else self.constant_color(hit[0].split(" ")[0], hit[1])
            mapname = ps.extract("string mapname")
            if mapname is not None:
                from PIL import Image
                img = np.asarray(Image.open(os.path.join(self.root, mapname)).convert("RGB"), np.uint8)
                self.sd.set_env_image(img, mult or (1.0, 1.0, 1.0))
            elif mult is not None:
                self.sd.set_env_constant(mult)
            else:
                raise PbrtError("can't process the infinite light")
            return
        pt = lambda name, default: (lambda h: default if h is None else tuple(float(v) for v in h[1]))(ps.extract_substr(name))
        if impl == "distant":
            frm, to = pt("from", (0.0, 0.0, 0.0)), pt("to", (0.0, 0.0, 1.0))
            L = self.color_arg(ps, "L", (1.0, 1.0, 1.0))
            d = tuple(float(F32(a) - F32(b)) for a, b in zip(to, frm))
            self.sd.add_distant_light(d, L, float("inf"))  # world radius fixed at commit, scene/src/lib.rs:54-58
        elif impl == "point":
            frm = pt("from", (0.0, 0.0, 0.0))
            self.sd.add_point_light(frm, self.color_arg(ps, "L", (1.0, 1.0, 1.0)))
        else:
            raise PbrtError(f"light of {impl} is unimplemented!() upstream")

    def world_item(self, item):  # :164-305
        kind = item[0]
        if kind == "Transform":
            self.ctm[-1] = self.ctm[-1] * self.transform_of(item[1])
        elif kind == "Shape":
            _, impl, ps = item
            ps.extract("alpha")
            ctm = self.ctm[-1]
            if self.area_l is not None and impl == "plymesh":
                # parse_samplable_shape, :408-433: one IsolatedTriangle instance + one triangle area
                # light (world-space vertices) per face
                P, _, _, idx = self.ply(ps)
                light_mtl = self.sd.diffuse_light(self.area_l)
                for i, j, k in idx:
                    tri = P[[i, j, k]]
                    w = [tuple(float(c) for c in ctm.apply_point(v)) for v in tri]  # transformed_by, sample_shape.rs:63-68
                    self.sd.add_area_light_triangle(w[0], w[1], w[2], self.area_l)
                    self.instances.append((self.sd.add_triangle(*(tuple(float(c) for c in v) for v in tri)), light_mtl, ctm))
            elif self.area_l is not None:
                if impl != "sphere":
                    raise PbrtError(f"samplable shape: {impl} is unimplemented!() upstream (only sphere / plymesh)")
                r = ps.lookup_f32("float radius")
                r = 1.0 if r is None else float(r)
                # SamplableShape::transformed_by, light/src/sample_shape.rs:46-82
                tx, ty, tz = ctm.apply_vec((1, 0, 0)), ctm.apply_vec((0, 1, 0)), ctm.apply_vec((0, 0, 1))
                cr = np.array([tx[1] * ty[2] - tx[2] * ty[1], tx[2] * ty[0] - tx[0] * ty[2], tx[0] * ty[1] - tx[1] * ty[0]], F32)
                scale = F32(np.cbrt(F32(F32(cr[0] * tz[0] + cr[1] * tz[1]) + cr[2] * tz[2])))
                if not scale > 0:
                    raise PbrtError("area-light transform must have positive uniform scale")
                center = ctm.apply_point((0.0, 0.0, 0.0))
                self.sd.add_area_light_sphere(tuple(float(c) for c in center), float(F32(r) * scale), self.area_l)
                light_mtl = self.sd.diffuse_light(self.area_l)
                self.instances.append((self.sd.add_sphere((0.0, 0.0, 0.0), r), light_mtl, ctm))
            elif self.mtl is not None:
                self.instances.append((self.shape(impl, ps), self.mtl, ctm))
            # else: "Neither arealight luminance or material are set" -> the shape is dropped (:196)
        elif kind == "Material":
            self.mtl = self.material(item[1], item[2])
        elif kind == "AttributeBlock":
            self.ctm.append(self.ctm[-1])
            self.mtl, self.area_l = None, None  # :203-204
            for child in item[1]:
                self.world_item(child)
            self.ctm.pop()
        elif kind == "TransformBlock":
            self.ctm.append(self.ctm[-1])
            for child in item[1]:
                self.world_item(child)
            self.ctm.pop()
        elif kind == "MakeNamedMaterial":
            ps = item[2]
            impl = ps.extract("string type")
            if not isinstance(impl, str):
                raise PbrtError("no material type specified")
            self.named_mtl[item[1]] = self.material(impl, ps)
        elif kind == "NamedMaterial":
            self.mtl = self.named_mtl.get(item[1])
        elif kind == "Texture":
            _, impl, ttype, name, ps = item
            if ttype in ("color", "spectrum"):
                if impl != "imagemap":
                    raise PbrtError(f"tex impl = {impl} (unimplemented!() upstream)")
                fn = ps.extract("string filename")
                if not isinstance(fn, str):
                    raise PbrtError("missing file name for image map texture")
                from PIL import Image
                img = np.asarray(Image.open(os.path.join(self.root, fn)).convert("RGB"), np.uint8)
                self.named_tex[name] = self.sd.add_texture_image(img)
        elif kind == "LightSource":
            self.light(item[1], item[2])
        elif kind == "AreaLightSource":
            if item[1] == "diffuse":
                hit = item[2].extract_substr("L")
                if hit is None or not isinstance(hit[1], list):
                    raise PbrtError("default / complicated luminance for diffuse light is unimplemented!() upstream")
                self.area_l = self.constant_color(hit[0].split(" ")[0], hit[1])
        # ReverseOrientation and anything else: "unhandled world item" upstream

    def load(self, options, items):
        # build_camera, :91-135
        fov = w = h = pose = None
        world = Affine()
        for opt in options:
            if opt[0] == "Camera":
                if opt[1] != "perspective":
                    pass  # logged upstream, then treated as perspective
                f = opt[2].extract("float fov")
                if isinstance(f, (list, str)):
                    raise PbrtError("complicated fov degree")
                fov = 60.0 if f is None else float(f)
            elif opt[0] == "Film":
                w, h = opt[2].lookup_f32("integer xresolution"), opt[2].lookup_f32("integer yresolution")
            elif opt[0] == "Transform" and opt[1][0] == "LookAt":
                pose = opt[1][1:]
        for opt in options:  # traverse_tree, :137-153: the remaining scene-wide transforms
            if opt[0] == "Transform" and opt[1][0] != "LookAt":
                world = world * self.transform_of(opt[1])
        if fov is None or w is None or h is None:
            raise PbrtError("the scene file needs Camera fov, Film xresolution and yresolution (the reference unwraps None)")
        eye, tgt, up = pose if pose else ((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0))
        self.sd.set_camera(int(w), int(h), fov, tuple(float(v) for v in eye), tuple(float(v) for v in tgt), tuple(float(v) for v in up))
        for it in items:
            self.world_item(it)
        for shape, mtl, ctm in self.instances:  # :158-160: world_transform * instance.transform
            t = world * ctm
            if t.is_identity():
                self.sd.add_instance(shape, mtl)
            else:
                self.sd.add_instance(shape, mtl, fwd=t.fwd, inv=t.inv)
        return self.sd


def load_pbrt_string(text, root_dir="."):
    options, items = Parser(tokenize(text, root_dir)).scene()
    return Loader(root_dir).load(options, items)


def load_pbrt(path):
    """scene::loader::build_scene(path) (scene/src/loader.rs:41-58) -> SceneDesc."""
    with open(path) as f:
        return load_pbrt_string(f.read(), os.path.dirname(os.path.abspath(path)))
