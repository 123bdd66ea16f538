"""Test-only environment that lets the reference's UNMODIFIED `exp_runner_blending.py` run end to end on top of the nudf
modules: none of this is on the product path.

  * `install_stubs()`   empty stand-ins for the third-party modules the runner imports but this image lacks and the training
                        loop never calls (trimesh, h5py, matplotlib, icecream, termcolor, mcubes, skimage, the Cython
                        `custom_mc`), and a small HOCON-subset reader registered as `pyhocon`;
  * `write_synthetic_dtu()`  a DTU/IDR-layout dataset directory (image/*.png, mask/*.png, cameras.npz with world_mat_i /
                        scale_mat_i) of an analytically rendered sphere, which `dataset/dataset.py` loads unchanged;
  * `write_conf()`      the reference's own conf text (confs/udf_dtu_blending.conf) with only paths / iteration counts /
                        frequencies replaced.
"""
import json
import math
import os
import re
import sys
import types

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# HOCON subset: nested `key { }` objects, `key = value`, lists, # and // comments, optional trailing commas
# ---------------------------------------------------------------------------------------------------------------
class ConfigTree(dict):
    def _walk(self, key, create=False):
        node = self
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node or not isinstance(dict.__getitem__(node, p), dict):
                if not create:
                    raise KeyError(key)
                dict.__setitem__(node, p, ConfigTree())
            node = dict.__getitem__(node, p)
        return node, parts[-1]

    def __getitem__(self, key):
        if isinstance(key, str) and "." in key:
            node, last = self._walk(key)
            return dict.__getitem__(node, last)
        return dict.__getitem__(self, key)

    def __setitem__(self, key, value):
        if isinstance(key, str) and "." in key:
            node, last = self._walk(key, create=True)
            dict.__setitem__(node, last, value)
        else:
            dict.__setitem__(self, key, value)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    _MISSING = object()

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def _typed(self, key, default, conv):
        try:
            v = self[key]
        except KeyError:
            if default is ConfigTree._MISSING:
                raise
            return default
        return conv(v)

    def get_string(self, key, default=_MISSING):
        return self._typed(key, default, str)

    def get_int(self, key, default=_MISSING):
        return self._typed(key, default, int)

    def get_float(self, key, default=_MISSING):
        return self._typed(key, default, float)

    def get_bool(self, key, default=_MISSING):
        return self._typed(key, default, lambda v: v if isinstance(v, bool) else str(v).lower() in ("true", "yes", "on", "1"))

    def get_list(self, key, default=_MISSING):
        return self._typed(key, default, list)

    def get_config(self, key, default=_MISSING):
        return self._typed(key, default, lambda v: v)


_TOKEN = re.compile(r'\s*(?:(#|//)[^\n]*|([{}\[\],=:])|"((?:[^"\\]|\\.)*)"|([^\s{}\[\],=:#"]+))')


def _tokens(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ValueError("HOCON-lite: cannot tokenise at %r" % text[pos:pos + 30])
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2):
            out.append(("p", m.group(2)))
        elif m.group(3) is not None:
            out.append(("s", m.group(3)))
        else:
            out.append(("w", m.group(4)))
    return out


def _scalar(tok):
    kind, v = tok
    if kind == "s":
        return v
    low = v.lower()
    if low in ("true", "false"):
        return low == "true"
    if low in ("null", "none"):
        return None
    try:
        return int(v)
    except ValueError:
        pass
    try:
        return float(v)
    except ValueError:
        return v


def _parse_value(toks, i):
    kind, v = toks[i]
    if (kind, v) == ("p", "{"):
        return _parse_object(toks, i + 1)
    if (kind, v) == ("p", "["):
        items, i = [], i + 1
        while toks[i] != ("p", "]"):
            if toks[i] == ("p", ","):
                i += 1
                continue
            val, i = _parse_value(toks, i)
            items.append(val)
        return items, i + 1
    return _scalar(toks[i]), i + 1


def _parse_object(toks, i, top=False):
    node = ConfigTree()
    while i < len(toks):
        if toks[i] == ("p", "}"):
            if top:
                raise ValueError("HOCON-lite: unbalanced }")
            return node, i + 1
        if toks[i] == ("p", ","):
            i += 1
            continue
        key = toks[i][1]
        i += 1
        if toks[i] in (("p", "="), ("p", ":")):
            i += 1
        val, i = _parse_value(toks, i)
        if isinstance(val, ConfigTree) and key in node and isinstance(node[key], ConfigTree):
            node[key].update(val)
        else:
            node[key] = val
    if not top:
        raise ValueError("HOCON-lite: missing }")
    return node, i


class ConfigFactory:
    @staticmethod
    def parse_string(text):
        tree, _ = _parse_object(_tokens(text), 0, top=True)
        return tree

    @staticmethod
    def parse_file(path):
        with open(path) as f:
            return ConfigFactory.parse_string(f.read())


class HOCONConverter:
    @staticmethod
    def to_hocon(conf, *a, **k):
        return json.dumps(conf, indent=2, default=str)


# ---------------------------------------------------------------------------------------------------------------
# stub modules
# ---------------------------------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _unavailable(what):
    def f(*a, **k):
        raise NotImplementedError("%s is a test stub (module absent from this image, not used by the training loop)" % what)
    return f


def install_stubs():
    here = sys.modules[__name__]
    _mod("pyhocon", ConfigFactory=ConfigFactory, HOCONConverter=HOCONConverter, ConfigTree=ConfigTree)
    _mod("icecream", ic=lambda *a, **k: (a[0] if a else None))
    _mod("termcolor", colored=lambda s, *a, **k: s)
    _mod("mcubes", marching_cubes=_unavailable("mcubes.marching_cubes"))
    sk = _mod("skimage")
    sk.measure = _mod("skimage.measure")
    tm = _mod("trimesh", Trimesh=_unavailable("trimesh.Trimesh"), load=_unavailable("trimesh.load"))
    tm.smoothing = _mod("trimesh.smoothing")
    _mod("h5py", File=_unavailable("h5py.File"))
    mpl = _mod("matplotlib")
    mpl.__path__ = []
    mpl.pyplot = _mod("matplotlib.pyplot", figure=_unavailable("matplotlib"), subplots=_unavailable("matplotlib"))
    def _get_cmap(name=None):
        def cmapper(value, bytes=False):             # a two-colour ramp standing in for matplotlib's colormaps (validate()'s
            v = np.clip(np.asarray(value, dtype=np.float64), 0.0, 1.0)     # depth visualisation, exp_runner_blending.py:847-866)
            rgba = np.stack([v, 0.2 + 0.6 * v, 1.0 - v, np.ones_like(v)], axis=-1)
            return (rgba * 255).astype(np.uint8) if bytes else rgba
        return cmapper
    mpl.cm = _mod("matplotlib.cm", get_cmap=_get_cmap)
    cm = _mod("custom_mc")
    cm.__path__ = []
    cm._marching_cubes_lewiner = _mod("custom_mc._marching_cubes_lewiner", udf_mc_lewiner=_unavailable("custom_mc"))
    return here


# ---------------------------------------------------------------------------------------------------------------
# synthetic DTU-layout dataset
# ---------------------------------------------------------------------------------------------------------------
def _look_at(eye):
    f = -eye / np.linalg.norm(eye)                              # camera looks at the origin
    up = np.array([0.0, 0.0, 1.0])
    r = np.cross(f, up); r /= np.linalg.norm(r)
    d = np.cross(f, r)                                          # image y axis points down
    R = np.stack([r, d, f], axis=0)                             # world -> camera
    t = -R @ eye
    return R, t


def write_synthetic_dtu(root, n_images=12, width=96, height=72, seed=0):
    """Sphere of radius 0.5 at the origin seen from a ring of cameras at distance 2.5 (outside the unit sphere, as
    dataset.py assumes); Lambert-ish colours from the normal; IDR camera file."""
    import cv2
    os.makedirs(os.path.join(root, "image"), exist_ok=True)
    os.makedirs(os.path.join(root, "mask"), exist_ok=True)
    rng = np.random.RandomState(seed)
    focal = 1.1 * width
    K = np.array([[focal, 0, (width - 1) / 2.0], [0, focal, (height - 1) / 2.0], [0, 0, 1.0]])
    cams = {}
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    for i in range(n_images):
        ang = 2 * math.pi * i / n_images
        eye = 2.5 * np.array([math.cos(ang), math.sin(ang), 0.35 + 0.1 * rng.rand()])
        eye = eye / np.linalg.norm(eye) * 2.5
        R, t = _look_at(eye)
        P = np.eye(4)
        P[:3, :4] = K @ np.concatenate([R, t[:, None]], axis=1)
        cams["world_mat_%d" % i] = P.astype(np.float64)
        cams["scale_mat_%d" % i] = np.eye(4)
        # ray-sphere intersection per pixel
        pix = np.stack([xs, ys, np.ones_like(xs)], axis=-1).reshape(-1, 3).astype(np.float64)
        dirs = (np.linalg.inv(K) @ pix.T).T
        dirs = dirs / np.linalg.norm(dirs, axis=1, keepdims=True)
        dirs = (R.T @ dirs.T).T
        b = dirs @ eye
        disc = b * b - (eye @ eye - 0.25)
        hit = disc > 0
        tt = -b - np.sqrt(np.maximum(disc, 0))
        pts = eye[None, :] + dirs * tt[:, None]
        nrm = pts / 0.5
        col = 0.5 + 0.5 * nrm
        shade = np.clip(0.3 + 0.7 * np.maximum(nrm @ (eye / 2.5), 0), 0, 1)[:, None]
        img = np.where(hit[:, None], col * shade, 0.0).reshape(height, width, 3)
        cv2.imwrite(os.path.join(root, "image", "%03d.png" % i), (img[:, :, ::-1] * 255).astype(np.uint8))
        cv2.imwrite(os.path.join(root, "mask", "%03d.png" % i), (hit.reshape(height, width, 1).repeat(3, 2) * 255).astype(np.uint8))
    np.savez(os.path.join(root, "cameras.npz"), **cams)
    return root


def write_conf(ref_root, out_path, data_dir, exp_dir, end_iter, batch_size=256, save_freq=2, val_freq=3,
               conf_name="confs/udf_dtu_blending.conf", extra_replace=()):
    """The reference's conf with paths / counts replaced (regex on `key = value` lines; everything else untouched)."""
    text = open(os.path.join(ref_root, conf_name)).read()

    def sub(key, val):
        nonlocal text
        text, n = re.subn(r"(?m)^(\s*%s\s*=\s*)[^\n#]*" % re.escape(key), lambda m: m.group(1) + str(val), text, count=1)
        assert n == 1, key
    sub("base_exp_dir", exp_dir)
    sub("data_dir", data_dir)
    sub("end_iter", end_iter)
    sub("batch_size", batch_size)
    sub("save_freq", save_freq)
    sub("val_freq", val_freq)
    sub("val_mesh_freq", 1000000000)
    sub("report_freq", 1)
    sub("validate_resolution_level", 4)
    sub("warm_up_end", 2)
    for k, v in extra_replace:
        sub(k, v)
    with open(out_path, "w") as f:
        f.write(text)
    return out_path
