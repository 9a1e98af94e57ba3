"""MJCF → flat kinematic tree ("model compiler", host side).

The reference hands the robot XML to MuJoCo (`mj.MjModel.from_xml_path`,
reference general_motion_retargeting/motion_retarget.py:27) and only ever uses the
kinematic subset of the compiled model: body tree, body pos/quat, hinge axes,
joint ranges/limited flags, the free joint, qpos0 and opt.timestep
(SURVEY.md §2 #7).  This module compiles exactly that subset without MuJoCo:

* ``<include file=…>`` expansion (nested; engineai_pm01/pm_v2.xml:1-4),
* ``<default>`` class inheritance + ``childclass`` for joint attributes
  (fourier_n1/n1_mocap.xml:24-31 puts ``range="0 0"`` in class defaults),
* ``<compiler angle autolimits eulerseq>``, ``<option timestep>``,
* ``<freejoint/>`` / ``<joint type="free"/>`` on the root body, 1-DoF hinges elsewhere.

The reference's own second statement of the same tree walk is
general_motion_retargeting/kinematics_model.py:101-163 (no include/default
support); tests compare body order/parents/offsets against it.

Body order is MuJoCo's: depth-first in document order, the world body dropped, so
body 0 is the floating root.  qpos = [x y z qw qx qy qz, hinge…] in body order.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

MJ_DEFAULT_TIMESTEP = 0.002  # MuJoCo's opt.timestep default


class MjcfError(ValueError):
    pass


@dataclass
class RobotModel:
    """Flat kinematic description of one floating-base robot."""

    name: str
    body_names: List[str]
    parent: np.ndarray        # [nbody] int32, -1 for the root
    body_pos: np.ndarray      # [nbody,3] float64, offset in the parent frame
    body_quat: np.ndarray     # [nbody,4] float64 wxyz, normalised
    body_hinge: np.ndarray    # [nbody] int32: hinge index owned by the body, -1 if none
    hinge_names: List[str]
    hinge_body: np.ndarray    # [nhinge] int32
    hinge_axis: np.ndarray    # [nhinge,3] float64, body-local, normalised
    hinge_lo: np.ndarray      # [nhinge] float64 (radians)
    hinge_hi: np.ndarray      # [nhinge] float64
    hinge_limited: np.ndarray  # [nhinge] bool
    qpos0: np.ndarray         # [7+nhinge] float64
    timestep: float = MJ_DEFAULT_TIMESTEP
    source: str = ""
    extras: Dict[str, object] = field(default_factory=dict)

    @property
    def nbody(self) -> int:
        return len(self.body_names)

    @property
    def nhinge(self) -> int:
        return len(self.hinge_names)

    @property
    def nq(self) -> int:
        return 7 + self.nhinge

    @property
    def nv(self) -> int:
        return 6 + self.nhinge

    def body_id(self, name: str) -> int:
        try:
            return self.body_names.index(name)
        except ValueError:
            raise KeyError(f"body '{name}' not in robot model '{self.name}'") from None

    def chain_hinges(self, body: int) -> List[int]:
        """Hinge indices on the path root→body, root side first."""
        out = []
        b = body
        while b >= 0:
            if self.body_hinge[b] >= 0:
                out.append(int(self.body_hinge[b]))
            b = int(self.parent[b])
        return out[::-1]

    # -- (de)serialisation to the plain-JSON "model pack" form -----------------------
    def to_dict(self) -> dict:
        return {
            "name": self.name,
            "body_names": list(self.body_names),
            "parent": self.parent.tolist(),
            "body_pos": self.body_pos.tolist(),
            "body_quat": self.body_quat.tolist(),
            "body_hinge": self.body_hinge.tolist(),
            "hinge_names": list(self.hinge_names),
            "hinge_body": self.hinge_body.tolist(),
            "hinge_axis": self.hinge_axis.tolist(),
            "hinge_lo": self.hinge_lo.tolist(),
            "hinge_hi": self.hinge_hi.tolist(),
            "hinge_limited": [bool(x) for x in self.hinge_limited],
            "qpos0": self.qpos0.tolist(),
            "timestep": self.timestep,
            "source": self.source,
        }

    @staticmethod
    def from_dict(d: dict) -> "RobotModel":
        return RobotModel(
            name=d["name"],
            body_names=list(d["body_names"]),
            parent=np.asarray(d["parent"], np.int32),
            body_pos=np.asarray(d["body_pos"], np.float64).reshape(-1, 3),
            body_quat=np.asarray(d["body_quat"], np.float64).reshape(-1, 4),
            body_hinge=np.asarray(d["body_hinge"], np.int32),
            hinge_names=list(d["hinge_names"]),
            hinge_body=np.asarray(d["hinge_body"], np.int32),
            hinge_axis=np.asarray(d["hinge_axis"], np.float64).reshape(-1, 3),
            hinge_lo=np.asarray(d["hinge_lo"], np.float64),
            hinge_hi=np.asarray(d["hinge_hi"], np.float64),
            hinge_limited=np.asarray(d["hinge_limited"], bool),
            qpos0=np.asarray(d["qpos0"], np.float64),
            timestep=float(d.get("timestep", MJ_DEFAULT_TIMESTEP)),
            source=d.get("source", ""),
        )


# --------------------------------------------------------------------------------------
# XML helpers
# --------------------------------------------------------------------------------------
def _floats(text: str, n: Optional[int] = None, what: str = "") -> np.ndarray:
    vals = np.array([float(t) for t in text.split()], dtype=np.float64)
    if n is not None and vals.size != n:
        raise MjcfError(f"expected {n} numbers for {what}, got '{text}'")
    return vals


def _expand_includes(elem: ET.Element, this_dir: str, main_dir: str, depth: int = 0) -> None:
    """Replace every <include file=…/> below `elem` by the children of the included
    file's root element, recursively.  MuJoCo resolves the path against the including
    file's directory (older releases: the main file's) — try both."""
    if depth > 16:
        raise MjcfError("include nesting too deep (cycle?)")
    i = 0
    while i < len(elem):
        child = elem[i]
        if child.tag == "include":
            fname = child.attrib.get("file")
            if fname is None:
                raise MjcfError("<include> without file attribute")
            cands = [os.path.join(this_dir, fname), os.path.join(main_dir, fname)]
            path = next((c for c in cands if os.path.isfile(c)), None)
            if path is None:
                raise MjcfError(f"included file '{fname}' not found (tried {cands})")
            inc_root = ET.parse(path).getroot()
            _expand_includes(inc_root, os.path.dirname(path), main_dir, depth + 1)
            elem.remove(child)
            for k, sub in enumerate(list(inc_root)):
                elem.insert(i + k, sub)
            i += len(inc_root)
        else:
            _expand_includes(child, this_dir, main_dir, depth)
            i += 1


class _Defaults:
    """<default> class tree restricted to the <joint> element's attributes."""

    def __init__(self, root: ET.Element):
        self.joint: Dict[str, Dict[str, str]] = {"main": {}}
        for top in root.findall("default"):
            self._walk(top, parent_attrs={}, is_top=True)

    def _walk(self, node: ET.Element, parent_attrs: Dict[str, str], is_top: bool) -> None:
        name = node.attrib.get("class", "main" if is_top else None)
        if name is None:
            raise MjcfError("nested <default> without class")
        attrs = dict(parent_attrs)
        if is_top and name == "main":
            attrs.update(self.joint.get("main", {}))
        for j in node.findall("joint"):
            attrs.update(j.attrib)
        self.joint[name] = attrs
        for sub in node.findall("default"):
            self._walk(sub, attrs, is_top=False)

    def joint_attrs(self, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.attrib.get("class", childclass or "main")
        if cls not in self.joint:
            raise MjcfError(f"unknown default class '{cls}'")
        out = dict(self.joint[cls])
        out.update(elem.attrib)
        return out


def _quat_normalize(q: np.ndarray) -> np.ndarray:
    n = float(np.linalg.norm(q))
    if n < 1e-15:
        raise MjcfError("zero quaternion")
    return q / n


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def _body_quat(attrib: Dict[str, str], angle_scale: float, eulerseq: str) -> np.ndarray:
    """Orientation of a <body>: quat (wxyz) or euler; other MJCF spellings are not used
    by any supported robot and are rejected rather than silently ignored."""
    for bad in ("axisangle", "xyaxes", "zaxis"):
        if bad in attrib:
            raise MjcfError(f"body orientation attribute '{bad}' is not supported")
    if "quat" in attrib:
        return _quat_normalize(_floats(attrib["quat"], 4, "quat"))
    if "euler" in attrib:
        e = _floats(attrib["euler"], 3, "euler") * angle_scale
        q = np.array([1.0, 0.0, 0.0, 0.0])
        for ang, ax in zip(e, eulerseq):
            half = 0.5 * ang
            r = np.array([math.cos(half), 0.0, 0.0, 0.0])
            r[1 + "xyz".index(ax.lower())] = math.sin(half)
            # lower case = intrinsic (post-multiply), upper case = extrinsic (pre-multiply)
            q = _quat_mul(q, r) if ax.islower() else _quat_mul(r, q)
        return _quat_normalize(q)
    return np.array([1.0, 0.0, 0.0, 0.0])


# --------------------------------------------------------------------------------------
# compiler
# --------------------------------------------------------------------------------------
def load_mjcf(path: str, name: Optional[str] = None) -> RobotModel:
    path = os.fspath(path)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    main_dir = os.path.dirname(os.path.abspath(path))
    root = ET.parse(path).getroot()
    if root.tag != "mujoco":
        raise MjcfError(f"{path}: root element is <{root.tag}>, expected <mujoco>")
    _expand_includes(root, main_dir, main_dir)

    # <compiler> / <option> may appear several times (includes); later ones win per attribute.
    comp: Dict[str, str] = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    opt: Dict[str, str] = {}
    for o in root.findall("option"):
        opt.update(o.attrib)
    angle = comp.get("angle", "degree")
    if angle not in ("degree", "radian"):
        raise MjcfError(f"compiler angle='{angle}'")
    angle_scale = 1.0 if angle == "radian" else math.pi / 180.0
    autolimits = comp.get("autolimits", "true") == "true"   # MuJoCo >= 3.0 default
    eulerseq = comp.get("eulerseq", "xyz")
    if comp.get("coordinate", "local") != "local":
        raise MjcfError("compiler coordinate='global' is not supported")
    timestep = float(opt.get("timestep", MJ_DEFAULT_TIMESTEP))

    defaults = _Defaults(root)

    roots = []
    for wb in root.findall("worldbody"):
        roots.extend(wb.findall("body"))
    if not roots:
        raise MjcfError(f"{path}: no <body> under <worldbody>")
    # the robot is the first top-level body that carries a free joint
    def _has_free(b: ET.Element) -> bool:
        if b.find("freejoint") is not None:
            return True
        return any(defaults.joint_attrs(j, None).get("type", "hinge") == "free" for j in b.findall("joint"))

    robot_roots = [b for b in roots if _has_free(b)]
    if len(robot_roots) != 1:
        raise MjcfError(f"{path}: expected exactly one floating-base body tree, found {len(robot_roots)}")

    body_names: List[str] = []
    parent: List[int] = []
    body_pos: List[np.ndarray] = []
    body_quat: List[np.ndarray] = []
    body_hinge: List[int] = []
    hinge_names: List[str] = []
    hinge_body: List[int] = []
    hinge_axis: List[np.ndarray] = []
    hinge_lo: List[float] = []
    hinge_hi: List[float] = []
    hinge_limited: List[bool] = []
    hinge_ref: List[float] = []

    def add_body(node: ET.Element, par: int, childclass: Optional[str]) -> None:
        idx = len(body_names)
        bname = node.attrib.get("name", f"body{idx}")
        childclass = node.attrib.get("childclass", childclass)
        body_names.append(bname)
        parent.append(par)
        body_pos.append(_floats(node.attrib.get("pos", "0 0 0"), 3, f"pos of {bname}"))
        body_quat.append(_body_quat(node.attrib, angle_scale, eulerseq))
        body_hinge.append(-1)

        joints = [defaults.joint_attrs(j, childclass) for j in node.findall("joint")]
        nfree = len(node.findall("freejoint")) + sum(1 for a in joints if a.get("type", "hinge") == "free")
        hinges = [a for a in joints if a.get("type", "hinge") != "free"]
        if par < 0:
            if nfree != 1 or hinges:
                raise MjcfError(f"root body '{bname}' must carry exactly one free joint and nothing else")
        else:
            if nfree:
                raise MjcfError(f"free joint on non-root body '{bname}'")
            if len(hinges) > 1:
                raise MjcfError(f"body '{bname}' has {len(hinges)} joints; at most one hinge per body is supported")
            for a in hinges:
                jtype = a.get("type", "hinge")
                if jtype != "hinge":
                    raise MjcfError(f"joint type '{jtype}' on body '{bname}' is not supported")
                jpos = _floats(a.get("pos", "0 0 0"), 3, "joint pos")
                if np.any(jpos != 0.0):
                    raise MjcfError(f"joint on '{bname}' has non-zero pos; off-centre hinges are not supported")
                axis = _floats(a.get("axis", "0 0 1"), 3, "joint axis")
                n = float(np.linalg.norm(axis))
                if n < 1e-15:
                    raise MjcfError(f"zero joint axis on '{bname}'")
                rng = _floats(a.get("range", "0 0"), 2, "joint range") * angle_scale
                lim_attr = a.get("limited", "auto")
                if lim_attr == "true":
                    limited = True
                elif lim_attr == "false":
                    limited = False
                else:  # auto
                    limited = bool(autolimits and rng[0] < rng[1])
                    if not autolimits and (rng[0] != 0.0 or rng[1] != 0.0):
                        raise MjcfError(f"joint on '{bname}': range without limited and autolimits=false")
                body_hinge[idx] = len(hinge_names)
                hinge_names.append(a.get("name", f"joint{len(hinge_names)}"))
                hinge_body.append(idx)
                hinge_axis.append(axis / n)
                hinge_lo.append(float(rng[0]))
                hinge_hi.append(float(rng[1]))
                hinge_limited.append(limited)
                hinge_ref.append(float(a.get("ref", "0")) * angle_scale)
        for child in node.findall("body"):
            add_body(child, idx, childclass)

    rr = robot_roots[0]
    add_body(rr, -1, None)
    if any(r != 0.0 for r in hinge_ref):
        raise MjcfError("hinge 'ref' != 0 is not supported")

    nh = len(hinge_names)
    qpos0 = np.zeros(7 + nh)
    qpos0[0:3] = body_pos[0]
    qpos0[3:7] = body_quat[0]
    return RobotModel(
        name=name or root.attrib.get("model", os.path.basename(path)),
        body_names=body_names,
        parent=np.asarray(parent, np.int32),
        body_pos=np.asarray(body_pos, np.float64).reshape(-1, 3),
        body_quat=np.asarray(body_quat, np.float64).reshape(-1, 4),
        body_hinge=np.asarray(body_hinge, np.int32),
        hinge_names=hinge_names,
        hinge_body=np.asarray(hinge_body, np.int32),
        hinge_axis=np.asarray(hinge_axis, np.float64).reshape(-1, 3),
        hinge_lo=np.asarray(hinge_lo, np.float64),
        hinge_hi=np.asarray(hinge_hi, np.float64),
        hinge_limited=np.asarray(hinge_limited, bool),
        qpos0=qpos0,
        timestep=timestep,
        source=os.path.basename(path),
    )
