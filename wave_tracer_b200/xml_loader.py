"""Scene XML front-end: reads the reference's scene files (the format of /root/reference/scenes/**.xml, loaded there by
src/scene/loader/loader.cpp + src/scene/loader/xml/loader.cpp) into the host scene model of scene.py, so that a file written
for wave_tracer renders through wtgpu_render unchanged.

What is covered is what the procedural BASELINE scene needs (scenes/diffraction_simple/double_slits.xml + bits/geometry.xml) and the
procedural parts of the others; anything else raises SceneXmlError naming the element -- nothing is skipped silently:

  <default name value> + -D overrides, `$name` substitution (loader.cpp:69-86, 369-411: textual, before any parsing)
  <include path>                     a fragment with several top-level elements
  expressions                        "( ... )" / bare arithmetic with + - * / ^ comparisons && || ! true false (the reference evaluates
                                     them with tinyexpr-plusplus); quantities "expr unit" with mm/cm/m/um/nm/km, deg/rad, K, Hz..GHz
  <integrator type=plt_path|plt_bdpt> integer max_depth, boolean FSD / russian_roulette / MIS / *_direct_sampling, string direction
  <sensor type=virtual_plane|perspective> boolean enabled / ray_trace_only, transform to_world, quantity extent / alpha / fov,
                                     integer samples, <film type=array> (width, height, rfilter_scale, <response type=monochromatic>)
  <emitter type=spot|point|directional|area>
  <bsdf type=twosided|diffuse|dielectric|surface_spm|composite|scale> (+ id / <ref id>), <surface_profile type=dirac|fractal|gaussian>
  <spectrum constant=|rgb=|blackbody=|ITU=|emitter=(illuminant)|type=discrete|composite>
  <shape type=rectangle|cube|sphere> point p/x/y, transform to_world, boolean enabled, nested bsdf / ref / area emitter

`rgb=` spectra (the reference upsamples RGB to a spectrum) are represented by an object that refuses to be evaluated: the microwave /
single-wavelength BASELINE scenes only carry them in composite bins the sensor never queries.
"""
import ast
import math
import os
import re
import xml.etree.ElementTree as ET

import numpy as np

from . import scene as S


class SceneXmlError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------ expressions and quantities
_UNITS = {
    "m": ("len", 1.0), "mm": ("len", 1e-3), "cm": ("len", 1e-2), "um": ("len", 1e-6), "µm": ("len", 1e-6), "nm": ("len", 1e-9), "km": ("len", 1e3),
    "°": ("ang", math.pi / 180), "deg": ("ang", math.pi / 180), "rad": ("ang", 1.0),
    "K": ("temp", 1.0),
    "Hz": ("freq", 1.0), "kHz": ("freq", 1e3), "MHz": ("freq", 1e6), "GHz": ("freq", 1e9), "THz": ("freq", 1e12),
}
_C0 = 2.99792458e8
_ALLOWED = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.BoolOp, ast.Compare, ast.Constant, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow, ast.Mod,
            ast.USub, ast.UAdd, ast.Not, ast.And, ast.Or, ast.Eq, ast.NotEq, ast.Lt, ast.LtE, ast.Gt, ast.GtE, ast.Call, ast.Name, ast.Load)
_FUNCS = {"sqrt": math.sqrt, "sin": math.sin, "cos": math.cos, "tan": math.tan, "abs": abs, "min": min, "max": max, "floor": math.floor, "ceil": math.ceil,
          "exp": math.exp, "log": math.log, "pow": pow, "pi": math.pi, "true": True, "false": False}


def evaluate(expr):
    """A value expression after `$` substitution: number, or arithmetic / boolean expression in tinyexpr syntax."""
    e = expr.strip()
    if e == "":
        raise SceneXmlError("empty expression")
    py = e.replace("&&", " and ").replace("||", " or ").replace("^", "**")
    py = re.sub(r"!(?!=)", " not ", py)
    py = re.sub(r"(?<![\w.])\.(\d)", r"0.\1", py)          # ".05" -> "0.05"
    py = re.sub(r"(?<![\w.])0+(\d)", r"\1", py)             # leading zeros are not octal
    try:
        tree = ast.parse(py, mode="eval")
    except SyntaxError as ex:
        raise SceneXmlError(f"cannot parse expression {expr!r}") from ex
    for n in ast.walk(tree):
        if not isinstance(n, _ALLOWED):
            raise SceneXmlError(f"unsupported construct in expression {expr!r}")
        if isinstance(n, ast.Name) and n.id not in _FUNCS:
            raise SceneXmlError(f"unknown identifier {n.id!r} in expression {expr!r}")
    return eval(compile(tree, "<scene-xml>", "eval"), {"__builtins__": {}}, dict(_FUNCS))


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(": depth += 1
        elif ch == ")": depth -= 1
        if ch == sep and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    out.append(cur)
    return [x.strip() for x in out]


_QRE = re.compile(r"^(?P<expr>.*?)(?P<unit>(?:[A-Za-zµ]+|°))?$", re.S)


def quantity(s, kind=None):
    """"($S-.0001) mm" -> SI float (metres / radians / kelvin / hertz).  kind: expected dimension or None (dimensionless allowed)."""
    s = s.strip()
    m = re.match(r"^(.*?)\s*([A-Za-zµ°]+)$", s, re.S)
    unit = None
    if m and m.group(2) in _UNITS and m.group(1).strip() != "":
        body, unit = m.group(1).strip(), m.group(2)
    else:
        body = s
    v = float(evaluate(body))
    if unit is None:
        if kind in ("len", "ang", "temp", "freq") and v != 0.0:
            raise SceneXmlError(f"quantity {s!r} needs a unit")
        return v
    dim, f = _UNITS[unit]
    if kind == "wavelength" and dim == "freq":
        return _C0 / (v * f)
    if kind == "wavelength":
        kind = "len"
    if kind is not None and dim != kind:
        raise SceneXmlError(f"quantity {s!r}: expected a {kind} unit")
    return v * f


def qvec(s, kind=None, n=None):
    v = [quantity(p, kind) for p in _split_top(s)]
    if n is not None and len(v) != n:
        raise SceneXmlError(f"expected {n} components in {s!r}")
    return v


def qrange(s, kind):
    a, b = s.split("..")
    return quantity(a, kind), quantity(b, kind)


def boolean(s):
    v = evaluate(s)
    return bool(v)


def integer(s):
    v = evaluate(s)
    return int(v)          # "$res/4": truncation, as the reference's integer parser does for an integral quotient


def complex_value(s):
    """"(1,100i)" / "1.5" / "(.2,3i)" (spectrum constant=...)."""
    t = s.strip()
    m = re.match(r"^\(\s*([^,]+)\s*,\s*([^)]+?)i\s*\)$", t)
    if m:
        return complex(float(evaluate(m.group(1))), float(evaluate(m.group(2))))
    return complex(float(evaluate(t)), 0.0)


# ------------------------------------------------------------------------------------------------ the document
class RGBSpectrum(S.Spectrum):
    """`rgb="r,g,b"`: the reference upsamples to a spectrum (src/spectrum/rgb.cpp); not restated -- evaluating it is an error."""
    def __init__(self, rgb): self.rgb = rgb
    def value(self, k):
        if np.size(k) == 0: return np.zeros(np.shape(k), np.complex128)
        raise SceneXmlError("rgb spectra are not supported at the wavenumbers this sensor queries")


class IlluminantSpectrum(S.Spectrum):
    """`emitter="D65"` etc. (CIE standard illuminants, tabulated over the visible range in the reference's data/): zero outside 300-830 nm, which is
    all a microwave sensor ever asks; evaluating it inside the visible range is an error (the tables are not restated)."""
    def __init__(self, name, scale=1.0): self.name, self.scale = name, scale
    def value(self, k):
        k = np.asarray(k, np.float64)
        kvis_lo, kvis_hi = S.wavelen_to_wavenum(830e-9), S.wavelen_to_wavenum(300e-9)
        if np.any((k >= kvis_lo) & (k <= kvis_hi)):
            raise SceneXmlError(f"illuminant {self.name} is not tabulated: only sensors outside the visible range can carry it")
        return np.zeros(k.shape, np.complex128)


def _subst(text, defines):
    def rep(m):
        name = m.group(1)
        if name not in defines:
            raise SceneXmlError(f'unknown define "${name}"')
        return defines[name]
    return re.sub(r"\$([A-Za-z0-9_]+)", rep, text)


def _parse_file(path):
    txt = open(path, encoding="utf-8").read()
    # pugixml (the reference's parser) accepts a bare `&` and `<` inside attribute values ("$a==true && $b==false", "$x<3"); expat does not
    txt = re.sub(r"&(?!(?:amp|lt|gt|quot|apos|#\d+|#x[0-9a-fA-F]+);)", "&amp;", txt)
    txt = re.sub(r'"[^"<>]*<[^"<>]*"', lambda m: m.group(0).replace("<", "&lt;"), txt)
    try:
        return ET.fromstring(txt)
    except ET.ParseError:
        body = re.sub(r"<\?xml[^>]*\?>", "", txt)             # an <include> fragment: several top-level elements
        return ET.fromstring("<fragment>" + body + "</fragment>")


class _Loader:
    def __init__(self, path, defines, missing_meshes="error", standins=None, bitmap_standin=.5):
        self.standin_specs, self.standins, self.bitmap_standin = dict(standins or {}), [], float(bitmap_standin)
        self.missing_meshes, self.skipped = missing_meshes, []
        self.dir = os.path.dirname(os.path.abspath(path))
        self.root = _parse_file(path)
        if self.root.tag != "scene":
            raise SceneXmlError("root element must be <scene>")
        self._expand_includes(self.root, self.dir)
        self.defines = {k: str(v) for k, v in (defines or {}).items()}
        for d in self.root.findall("default"):
            self.defines.setdefault(d.attrib["name"], d.attrib["value"])
        for el in self.root.iter():
            for k, v in list(el.attrib.items()):
                if "$" in v:
                    el.attrib[k] = _subst(v, self.defines)
        self.bsdfs = {}

    def _expand_includes(self, node, base):
        for i, ch in enumerate(list(node)):
            if ch.tag == "include":
                p = os.path.join(base, ch.attrib["path"])
                frag = _parse_file(p)
                self._expand_includes(frag, os.path.dirname(p))
                idx = list(node).index(ch)
                node.remove(ch)
                for j, sub in enumerate(list(frag) if frag.tag == "fragment" else [frag]):
                    node.insert(idx + j, sub)
            else:
                self._expand_includes(ch, base)

    # ---- helpers
    @staticmethod
    def _point(p):
        """<point name=... x= y= z=/> or <point name=... value="x, y, z"/> (lengths)."""
        if "value" in p.attrib:
            return [float(v) for v in qvec(p.attrib["value"], "len", 3)]
        return [quantity(p.attrib.get(a, "0"), "len") for a in "xyz"]

    @staticmethod
    def _named(node, tag, name):
        for ch in node.findall(tag):
            if ch.attrib.get("name") == name:
                return ch
        return None

    def _enabled(self, node):
        b = self._named(node, "boolean", "enabled")
        return True if b is None else boolean(b.attrib["value"])

    def _bool(self, node, name, default):
        b = self._named(node, "boolean", name)
        return default if b is None else boolean(b.attrib["value"])

    def _int(self, node, name, default):
        b = self._named(node, "integer", name)
        return default if b is None else integer(b.attrib["value"])

    def _float(self, node, name, default):
        b = self._named(node, "float", name)
        return default if b is None else float(evaluate(b.attrib["value"]))

    def _quantity(self, node, name, kind, default=None):
        b = self._named(node, "quantity", name)
        return default if b is None else quantity(b.attrib["value"], kind)

    def transform(self, node):
        """src/math/transform_loader.cpp:60-125."""
        if node is None:
            return np.eye(4)
        la = node.find("lookat")
        if la is not None:
            origin = np.array(qvec(la.attrib["origin"], "len", 3)) if "origin" in la.attrib else np.zeros(3)
            target = np.array(qvec(la.attrib["target"], "len", 3)) if "target" in la.attrib else np.array([0, 0, 1.0])
            d = target - origin; d = d / np.linalg.norm(d)
            if "up" in la.attrib:
                up = np.array(qvec(la.attrib["up"], None, 3), np.float64)
            else:       # frame_t::build_orthogonal_frame(dir).t  (include/wt/math/frame.hpp:158-174)
                n = d
                if abs(n[0]) > abs(n[1]): x = 1 / math.sqrt(n[0] ** 2 + n[2] ** 2); b = np.array([x * n[2], 0, -x * n[0]])
                else: x = 1 / math.sqrt(n[1] ** 2 + n[2] ** 2); b = np.array([0, x * n[2], -x * n[1]])
                up = np.cross(b, n) + 0.0          # (+0.0: no negative zeros in the frame)
            if 1 - abs(float(np.dot(up / np.linalg.norm(up), d))) < 1e-5:
                raise SceneXmlError("degenerate 'lookat' transform")
            return S.lookat(origin, target, up)
        M = np.eye(4)
        for it in node:
            if it.tag == "translate":
                M = S.translate(self._point(it)) @ M
            elif it.tag == "scale":
                if "value" in it.attrib: v = float(evaluate(it.attrib["value"])); M = S.scale((v, v, v)) @ M
                else: M = S.scale([float(evaluate(it.attrib.get(a, "1"))) for a in "xyz"]) @ M
            elif it.tag == "rotate":
                M = S.rotate([float(evaluate(it.attrib.get(a, "0"))) for a in "xyz"], quantity(it.attrib["angle"], "ang")) @ M
            elif it.tag == "matrix":
                vals = [quantity(x, "len") if re.search(r"[a-zA-Z]\s*$", x) else float(evaluate(x)) for x in _split_top(it.attrib["value"].strip())]       # "0, 1, 0, 0cm, ..."
                M = np.array(vals, np.float64).reshape(4, 4) @ M
            else:
                raise SceneXmlError(f"<transform>: unsupported child <{it.tag}>")
        return M

    # ---- spectra
    def spectrum(self, node):
        a = node.attrib
        scale = self._float(node, "scale", 1.0)
        if "constant" in a:
            v = complex_value(a["constant"]) * scale
            return S.Const(v)
        if "rgb" in a:
            return RGBSpectrum(qvec(a["rgb"], None, 3))
        if "blackbody" in a:
            return S.Blackbody(quantity(a["blackbody"], "temp"), scale)
        if "ITU" in a:
            if a["ITU"] not in S.ITU.TABLE: raise SceneXmlError(f'<spectrum ITU="{a["ITU"]}">: unknown ITU-R P.2040 material')
            if scale != 1.0: raise SceneXmlError("scaled ITU spectra are not supported")
            return S.ITU(a["ITU"])
        if "material" in a:        # refractive index database (data/ior/<name>.yml): complex IOR
            try: m = S.Material(a["material"])
            except ValueError as ex: raise SceneXmlError(str(ex)) from ex
            return m if scale == 1.0 else S.Scaled(m, scale)
        if "emitter" in a:
            if f'emission/{a["emitter"]}/value' in S.spectra_db(): return S.Emission(a["emitter"], scale)
            return IlluminantSpectrum(a["emitter"], scale)
        t = a.get("type")
        if t == "discrete":
            return S.Discrete(quantity(a["wavelength"], "wavelength"), float(evaluate(a.get("value", "1"))) * scale)
        if t == "composite":
            bins = []
            for b in node.findall("bin"):
                lo, hi = qrange(b.attrib["wavelength_range"], "wavelength")
                sub = b.find("spectrum")
                if sub is None: raise SceneXmlError("composite spectrum: <bin> without <spectrum>")
                bins.append((lo, hi, self.spectrum(sub)))
            return S.Binned(bins)
        raise SceneXmlError(f"<spectrum>: unsupported form {dict(a)}")

    def texture(self, node):
        """texture_t (include/wt/texture/texture.hpp) reduced to what the device evaluates: a spectrum.  constant and scale textures are exact;
        a bitmap is replaced by its stand-in mean value (the reference's PNG / EXR files are Git-LFS stubs): `bitmap_standin` (default 0.5)."""
        t = node.attrib.get("type", "constant")
        if t == "constant":
            sp = node.find("spectrum")
            return self.spectrum(sp) if sp is not None else S.Const(float(evaluate(node.attrib.get("value", "1"))))
        if t == "scale":
            sc = self._spectrum_child(node, "scale", S.Const(1.0))
            inner = [self.texture(ch) for ch in node.findall("texture")]
            if len(inner) != 1: raise SceneXmlError("scale texture needs exactly one nested texture")
            if not isinstance(sc, S.Const): raise SceneXmlError("scale texture: only constant scales are supported")
            return S.Scaled(inner[0], sc.v.real)
        if t == "bitmap":
            pth = node.find("path")
            self.standins.append(("texture", pth.attrib.get("value") if pth is not None else "?", f"constant {self.bitmap_standin}"))
            return S.Const(self.bitmap_standin)
        raise SceneXmlError(f"<texture type={t!r}> is not supported")

    def _spectrum_child(self, node, name, default=None):
        ch = self._named(node, "spectrum", name)
        if ch is None:
            tx = self._named(node, "texture", name)
            if tx is not None: return self.texture(tx)
        return default if ch is None else self.spectrum(ch)

    # ---- bsdfs
    def surface_profile(self, node):
        if node is None:
            return S.Dirac()
        t = node.attrib.get("type")
        if t == "dirac":
            return S.Dirac()
        if t == "fractal":
            r = self._spectrum_child(node, "roughness")
            if r is None: raise SceneXmlError("fractal surface_profile: only the roughness parametrisation is supported")
            return S.Fractal(r, gamma=self._float(node, "gamma", 3.0))
        if t == "gaussian":
            r = self._spectrum_child(node, "roughness")
            if r is not None: return S.Gaussian(roughness=r)
            q = self._named(node, "quantity", "sigma")
            if q is None: raise SceneXmlError("gaussian surface_profile: either 'roughness' or 'sigma' must be provided")
            # rms_t is 1/mm (surface_profile.hpp:41): "<value> 1/mm" is not a unit this loader parses; accept a bare number in 1/mm
            return S.Gaussian(sigma=float(evaluate(q.attrib["value"].replace("1/mm", "").replace("/mm", ""))))
        raise SceneXmlError(f"<surface_profile type={t!r}> is not supported")

    def bsdf(self, node):
        t = node.attrib.get("type")
        nested = [self.bsdf(ch) for ch in node.findall("bsdf")] + [self._ref(ch) for ch in node.findall("ref")]
        if t is None and "scale" in node.attrib:       # <bsdf scale=".1"> ... </bsdf>  (src/bsdf/bsdf_loader.cpp:36-53)
            if len(nested) != 1: raise SceneXmlError("scale bsdf needs exactly one nested bsdf")
            out = S.Scale(S.Const(float(evaluate(node.attrib["scale"]))), nested[0])
        elif t == "twosided":
            if len(nested) != 1: raise SceneXmlError("twosided bsdf needs exactly one nested bsdf")
            out = S.TwoSided(nested[0])
        elif t == "diffuse":
            out = S.Diffuse(self._spectrum_child(node, "reflectance", S.Const(.5)))
        elif t in ("dielectric", "surface_spm"):
            ior = self._spectrum_child(node, "IOR")
            if ior is None: raise SceneXmlError(f"{t} bsdf: 'IOR' spectrum must be provided")
            ext = self._spectrum_child(node, "extIOR", S.Const(1.0))
            rs, ts = self._spectrum_child(node, "reflection_scale"), self._spectrum_child(node, "transmission_scale")
            if t == "dielectric": out = S.Dielectric(ior, ext, rs, ts)
            else: out = S.SurfaceSPM(ior, ext, self.surface_profile(node.find("surface_profile")), rs, ts)
        elif t == "composite":
            bins = []
            for b in node.findall("bin"):
                lo, hi = qrange(b.attrib["wavelength_range"], "wavelength")
                subs = [self.bsdf(ch) for ch in b.findall("bsdf")] + [self._ref(ch) for ch in b.findall("ref")]
                if len(subs) != 1: raise SceneXmlError("composite bsdf: each <bin> needs exactly one bsdf")
                bins.append((lo, hi, subs[0]))
            out = S.Composite(bins)
        elif t == "scale":
            if len(nested) != 1: raise SceneXmlError("scale bsdf needs exactly one nested bsdf")
            out = S.Scale(self._spectrum_child(node, "scale", S.Const(1.0)), nested[0])
        else:
            raise SceneXmlError(f"<bsdf type={t!r}> is not supported")
        if "id" in node.attrib:
            self.bsdfs[node.attrib["id"]] = out
        return out

    def _ref(self, node):
        i = node.attrib["id"]
        if i not in self.bsdfs:
            raise SceneXmlError(f'<ref id="{i}">: unknown id')
        return self.bsdfs[i]

    # ---- integrator / sensor / emitters / shapes
    def integrator(self, node, lut):
        t = node.attrib.get("type")
        md = self._int(node, "max_depth", 1024)
        fsd = self._bool(node, "FSD", True)
        rr = self._bool(node, "russian_roulette", True)
        if t == "plt_path":
            d = self._named(node, "string", "direction")
            return S.PltPath(max_depth=md, direction=d.attrib["value"] if d is not None else "backward", fsd=fsd, russian_roulette=rr)
        if t == "plt_bdpt":
            return S.PltBdpt(max_depth=md, fsd=fsd, russian_roulette=rr, mis=self._bool(node, "MIS", True),
                             sensor_direct_sampling=self._bool(node, "sensor_direct_sampling", True),
                             emitter_direct_sampling=self._bool(node, "emitter_direct_sampling", True), lut=lut)
        raise SceneXmlError(f"<integrator type={t!r}> is not supported")

    def film(self, node):
        if node is None or node.attrib.get("type") != "array":
            raise SceneXmlError("sensor needs a <film type=\"array\">")
        resp = node.find("response")
        rt = resp.attrib.get("type") if resp is not None else None
        if rt == "monochromatic":
            channels = [self.spectrum(resp.find("spectrum"))]
        elif rt == "RGB":       # src/sensor/response/RGB.cpp: defaults sRGB / D65
            cs = self._named(resp, "string", "colourspace"); wp = self._named(resp, "string", "white_point")
            try: channels = S.rgb_response(cs.attrib["value"] if cs is not None else "sRGB", wp.attrib["value"] if wp is not None else "D65")
            except KeyError as ex: raise SceneXmlError(f"RGB response: unsupported colourspace / white point {ex}") from ex
        else:
            raise SceneXmlError(f"<response type={rt!r}> is not supported (monochromatic, RGB)")
        return S.Film(self._int(node, "width", 0), self._int(node, "height", 0), channels, rfilter_scale=self._float(node, "rfilter_scale", 1.0))

    def sensor(self, node):
        t = node.attrib.get("type")
        tw = self.transform(self._named(node, "transform", "to_world"))
        film = self.film(node.find("film"))
        spp = self._int(node, "samples", 1)
        rt = self._bool(node, "ray_trace_only", False)
        if t == "virtual_plane":
            ext = self._named(node, "quantity", "extent")
            if ext is None: raise SceneXmlError("virtual_plane sensor: 'extent' must be provided")
            return S.VirtualPlane(tw, tuple(qvec(ext.attrib["value"], "len", 2)), film, alpha=self._quantity(node, "alpha", "ang"), ray_trace_only=rt, samples=spp)
        if t == "perspective":
            return S.Perspective(tw, self._quantity(node, "fov", "ang"), film, ray_trace_only=rt, samples=spp)
        raise SceneXmlError(f"<sensor type={t!r}> is not supported")

    def emitter(self, node):
        t = node.attrib.get("type")
        tw_node = self._named(node, "transform", "to_world")
        pse = self._float(node, "phase_space_extent_scale", 1.0)
        if t == "spot":
            kw = {}
            c = self._quantity(node, "cutoff_angle", "ang"); b = self._quantity(node, "beam_width", "ang")
            if c is not None: kw["cutoff_angle"] = c
            if b is not None: kw["beam_width"] = b
            return S.Spot(self.transform(tw_node), self._spectrum_child(node, "radiant_intensity"), phase_space_extent_scale=pse, **kw)
        if t == "point":
            p = self._named(node, "point", "position")
            pos = self._point(p) if p is not None else [0, 0, 0]
            return S.Point(pos, self._spectrum_child(node, "radiant_intensity"), phase_space_extent_scale=pse)
        if t == "directional":
            return S.Directional(self._spectrum_child(node, "irradiance"), self.transform(tw_node) if tw_node is not None else None, phase_space_extent_scale=pse)
        if t == "area":
            return S.Area(self._spectrum_child(node, "radiance"), scale=self._float(node, "scale", 1.0), phase_space_extent_scale=pse)
        raise SceneXmlError(f"<emitter type={t!r}> is not supported")

    def shape(self, node):
        t = node.attrib.get("type")
        tw_node = self._named(node, "transform", "to_world")
        tw = self.transform(tw_node) if tw_node is not None else None
        def pt(name):
            p = self._named(node, "point", name)
            if p is None: raise SceneXmlError(f"{t} shape: point '{name}' must be provided")
            return np.array(self._point(p))
        def centre_of(name="center"):
            c = self._named(node, "point", name)
            return self._point(c) if c is not None else (0, 0, 0)
        if t == "rectangle":        # src/scene/shape.cpp:196-227
            ln = self._quantity(node, "length", "len")
            tess = self._int(node, "tessellation", 1)
            mesh = S.square(ln, to_world=tw, tessellation=tess) if ln is not None else S.rectangle(pt("p"), pt("x"), pt("y"), to_world=tw, tessellation=tess)
        elif t == "cube":
            mesh = S.cube_len(self._quantity(node, "length", "len", 2.0), tw)
        elif t == "prism":
            mesh = S.prism(self._quantity(node, "length", "len", 1.0), self._quantity(node, "height", "len", 1.0), self._quantity(node, "angle", "ang", math.pi / 2), to_world=tw)
        elif t == "sphere":
            mesh = S.icosphere(self._quantity(node, "radius", "len", 1.0), centre_of(), self._int(node, "tessellation", 32), to_world=tw)
        elif t == "cylinder":
            mesh = S.cylinder(pt("p0"), pt("p1"), self._quantity(node, "radius", "len", 1.0), self._int(node, "tessellation", 32), to_world=tw)
        elif t == "lens":
            mesh = S.lens(self._quantity(node, "radius", "len", 1.0), centre_of(), self._float(node, "R1", 0.0), self._float(node, "R2", 0.0),
                          self._quantity(node, "thickness", "len", 0.0), self._int(node, "tessellation", 50), to_world=tw)
        elif t in ("ply", "obj") and self.missing_meshes == "standin":
            # SYNTHETIC: the mesh file is a Git-LFS stub; a procedural closed surface with the triangle budget and placement the scene gives it
            # (standins: {shape id: dict(kind="blob"|"star", tris=..., radius=..., centre=...)}) takes its place and is listed in scene.standins
            sid = node.attrib.get("id", "?")
            spec = self.standin_specs.get(sid)
            if spec is None: raise SceneXmlError(f'<shape type="{t}" id="{sid}">: no stand-in given for this mesh')
            unit = self._quantity(node, "scale", "len", 1.0)
            M = (tw if tw is not None else np.eye(4)) @ S.scale(unit)
            if spec["kind"] == "blob":      # one or several lobes: (radius, centre, triangles)
                lobes = spec.get("lobes") or [(spec["radius"], spec.get("centre", (0, 0, 0)), spec["tris"])]
                parts = [S.blob(r, c, n, seed=spec.get("seed", 1) + 7 * i) for i, (r, c, n) in enumerate(lobes)]
                pos = np.concatenate([m.positions for m in parts]); nrm = np.concatenate([m.normals for m in parts])
                off = np.cumsum([0] + [len(m.positions) for m in parts[:-1]])
                mesh = S.Mesh(pos, np.concatenate([m.indices + o for m, o in zip(parts, off)]), normals=nrm, to_world=M)
            else: mesh = S.star_prism(spec.get("outer", 1.0), spec.get("inner", .55), spec.get("depth", 1.0), to_world=M)
            fn = node.find("path")
            self.standins.append(("mesh", fn.attrib.get("value") if fn is not None else sid, f'{spec["kind"]} of {len(mesh.indices)} triangles'))
        elif t in ("ply", "obj") and self.missing_meshes == "skip":
            fn = self._named(node, "string", "filename") or self._named(node, "path", "filename")
            self.skipped.append((node.attrib.get("id", "?"), fn.attrib.get("value") if fn is not None else "?"))
            for ch in node.findall("bsdf"): self.bsdf(ch)          # nested materials may carry ids other shapes refer to
            return None
        else:
            raise SceneXmlError(f"<shape type={t!r}> is not supported (ply/obj meshes are Git-LFS stubs in the reference tree; "
                                "missing_meshes=\"skip\" loads the rest of the scene and lists them)")
        bs = [self.bsdf(ch) for ch in node.findall("bsdf")] + [self._ref(ch) for ch in node.findall("ref")]
        if len(bs) != 1: raise SceneXmlError("a shape needs exactly one bsdf (nested or <ref>)")
        em = node.find("emitter")
        return mesh, bs[0], (self.emitter(em) if em is not None else None)

    def build(self, lut=(2048, 1024), sensor_id=None):
        sc = S.Scene()
        root = self.root
        integs = [i for i in root.findall("integrator") if self._enabled(i)]
        if len(integs) != 1: raise SceneXmlError(f"{len(integs)} enabled <integrator> elements; exactly one is needed")
        sc.integrator = self.integrator(integs[0], lut)
        sensors = [s for s in root.findall("sensor") if self._enabled(s) and (sensor_id is None or s.attrib.get("id") == sensor_id)]
        if len(sensors) != 1:
            raise SceneXmlError(f"{len(sensors)} enabled sensors; exactly one is rendered per call (select with sensor_id)")
        sc.sensor = self.sensor(sensors[0])
        for b in root.findall("bsdf"):
            self.bsdf(b)
        for e in root.findall("emitter"):
            if self._enabled(e): sc.add_emitter(self.emitter(e))
        for sh in root.findall("shape"):
            if not self._enabled(sh): continue
            r = self.shape(sh)
            if r is None: continue
            mesh, bsdf, em = r
            sc.add_shape(mesh, bsdf, emitter=em)
        known = {"default", "integrator", "sensor", "bsdf", "emitter", "shape", "sampler"}
        for ch in root:
            if ch.tag not in known:
                raise SceneXmlError(f"unsupported top-level element <{ch.tag}>")
        smp = root.find("sampler")
        if smp is not None:
            t = smp.attrib.get("type")
            if t in ("sobolld", "sobol"): sc.sampler = S.Sobolld()
            elif t not in ("uniform", "independent"): raise SceneXmlError(f"<sampler type={t!r}> is not supported")
        sc.xml_bsdfs = dict(self.bsdfs)             # materials by id, as <ref> resolves them
        sc.skipped_shapes = list(self.skipped)      # (id, file) of mesh shapes left out under missing_meshes="skip"
        sc.standins = list(self.standins)           # (kind, file, what replaced it): every synthetic substitution made while loading
        return sc


def load_scene(path, defines=None, lut=(2048, 1024), sensor_id=None, missing_meshes="error", standins=None, bitmap_standin=.5):
    """`wave_tracer render scene.xml -D k=v,...` front half: returns a scene.Scene (call .build() for the wtgpu_scene_desc tables).
    missing_meshes="skip": ply/obj shapes (Git-LFS stubs in the reference tree) are left out and listed in scene.skipped_shapes;
    missing_meshes="standin": they are replaced by the procedural surfaces `standins` describes per shape id (listed in scene.standins)."""
    return _Loader(path, defines, missing_meshes, standins, bitmap_standin).build(lut=lut, sensor_id=sensor_id)


def parse_defines(s):
    """"res=1440,spp=1024" -> dict (the CLI's -D option, src/main.cpp)."""
    out = {}
    for kv in (s or "").split(","):
        if kv.strip():
            k, v = kv.split("=", 1); out[k.strip()] = v.strip()
    return out
