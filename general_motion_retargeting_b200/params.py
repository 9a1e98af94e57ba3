"""Robot / IK-config registry and compiled "model packs".

Same keys as the reference registry (reference general_motion_retargeting/params.py:7-63:
``ROBOT_XML_DICT``, ``IK_CONFIG_DICT``, ``ROBOT_BASE_DICT``,
``VIEWER_CAM_DISTANCE_DICT``).  The reference resolves them to files inside its own
checkout; here every (source, robot) pair is also shipped as a compiled *model pack*
(``data/packs/<src>_to_<robot>.json`` = flat kinematic tree + parsed IK config, produced
by ``tools/compile_packs.py`` from the reference's MJCF/JSON) so that the solver runs
where the reference checkout does not exist.  When a reference checkout is available
(``GMR_REFERENCE_ROOT`` or ``/root/reference``) the raw files can be compiled on the fly
with ``compile_pack``; tests check both routes agree.
"""
from __future__ import annotations

import hashlib
import json
import os
import pathlib
import re
import warnings
from typing import Dict, Optional, Tuple

from .ik_config import IKConfig
from .mjcf import RobotModel, load_mjcf

HERE = pathlib.Path(__file__).parent
PACK_ROOT = HERE / "data" / "packs"

# robot key -> MJCF path relative to the reference's assets/ directory
ROBOT_XML_REL = {
    "unitree_g1": "unitree_g1/g1_mocap_29dof.xml",
    "booster_t1": "booster_t1/t1_mocap.xml",
    "booster_t1_4dof": "booster_t1/t1_mocap_4dof.xml",
    "stanford_toddy": "stanford_toddy/toddy_mocap.xml",
    "fourier_n1": "fourier_n1/n1_mocap.xml",
    "engineai_pm01": "engineai_pm01/pm_v2.xml",
    "kuavo_s45": "kuavo_s45/biped_s45_collision.xml",
    "hightorque_hi": "hightorque_hi/hi_25dof.xml",
}

# source format -> robot key -> file name under the reference's ik_configs/
IK_CONFIG_REL = {
    "smplx": {
        "unitree_g1": "smplx_to_g1.json",
        "booster_t1": "smplx_to_t1.json",
        "stanford_toddy": "smplx_to_toddy.json",
        "fourier_n1": "smplx_to_n1.json",
        "engineai_pm01": "smplx_to_pm01.json",
        "kuavo_s45": "smplx_to_kuavo.json",
        "hightorque_hi": "smplx_to_hi.json",
    },
    "bvh": {
        "unitree_g1": "bvh_to_g1.json",
        "booster_t1": "bvh_to_t1.json",
        "booster_t1_4dof": "bvh_to_t1_4dof.json",
        "fourier_n1": "bvh_to_n1.json",
        "stanford_toddy": "bvh_to_toddy.json",
        "engineai_pm01": "bvh_to_pm01.json",
    },
    "fbx": {
        "unitree_g1": "fbx_to_g1.json",
    },
}

ROBOT_BASE_DICT = {
    "unitree_g1": "pelvis",
    "booster_t1": "Waist",
    "booster_t1_4dof": "Waist",
    "stanford_toddy": "waist_link",
    "fourier_n1": "base_link",
    "engineai_pm01": "LINK_BASE",
    "kuavo_s45": "base_link",
    "hightorque_hi": "base_link",
}

VIEWER_CAM_DISTANCE_DICT = {k: (1.0 if k == "stanford_toddy" else 2.0) for k in ROBOT_XML_REL}


def reference_root() -> Optional[pathlib.Path]:
    """Directory of a reference checkout (holding assets/ and
    general_motion_retargeting/ik_configs/), or None."""
    cands = []
    if os.environ.get("GMR_REFERENCE_ROOT"):
        cands.append(pathlib.Path(os.environ["GMR_REFERENCE_ROOT"]))
    cands.append(pathlib.Path("/root/reference"))
    for c in cands:
        if (c / "assets").is_dir() and (c / "general_motion_retargeting" / "ik_configs").is_dir():
            return c
    return None


def _paths(root: Optional[pathlib.Path]):
    asset_root = (root / "assets") if root else pathlib.Path("assets")
    ik_root = (root / "general_motion_retargeting" / "ik_configs") if root else pathlib.Path("ik_configs")
    return asset_root, ik_root


def _build_dicts():
    asset_root, ik_root = _paths(reference_root())
    xml = {k: asset_root / v for k, v in ROBOT_XML_REL.items()}
    ik = {s: {r: ik_root / f for r, f in d.items()} for s, d in IK_CONFIG_REL.items()}
    return xml, ik


ROBOT_XML_DICT, IK_CONFIG_DICT = _build_dicts()


def pack_path(src_human: str, tgt_robot: str) -> pathlib.Path:
    return PACK_ROOT / f"{src_human}_to_{tgt_robot}.json"


_INCLUDE_RE = re.compile(rb"<include\s+file\s*=\s*[\"']([^\"']+)[\"']")


def source_digest(xml_path: os.PathLike, ik_path: os.PathLike) -> str:
    """sha256 over the bytes of the MJCF, every file it <include>s (recursively, in order) and the IK JSON:
    the identity of the SOURCES a pack was compiled from.  `load_pack` compares it with the checkout's files so
    that an edited offset, weight, scale or joint range is never silently ignored in favour of a stale pack."""
    h = hashlib.sha256()
    main_dir = os.path.dirname(os.fspath(xml_path))

    def feed(path, depth=0):
        with open(path, "rb") as f:
            data = f.read()
        h.update(data)
        if depth < 8:
            for m in _INCLUDE_RE.finditer(data):
                name = m.group(1).decode()
                for cand in (os.path.join(os.path.dirname(path), name), os.path.join(main_dir, name)):
                    if os.path.isfile(cand):
                        feed(cand, depth + 1)
                        break

    feed(os.fspath(xml_path))
    h.update(b"\0ik\0")
    with open(ik_path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def compile_pack(src_human: str, tgt_robot: str, root: Optional[os.PathLike] = None) -> dict:
    """Compile the raw MJCF + IK JSON of a reference checkout into a pack dict."""
    rroot = pathlib.Path(root) if root is not None else reference_root()
    if rroot is None:
        raise FileNotFoundError("no reference checkout found (set GMR_REFERENCE_ROOT)")
    asset_root, ik_root = _paths(rroot)
    cfg_file = IK_CONFIG_REL[src_human][tgt_robot]       # KeyError like the reference (:30)
    xml_rel = ROBOT_XML_REL[tgt_robot]
    robot = load_mjcf(asset_root / xml_rel, name=tgt_robot)
    cfg = IKConfig.from_json(ik_root / cfg_file)
    return {
        "src_human": src_human,
        "tgt_robot": tgt_robot,
        "xml_rel": xml_rel,
        "ik_config_file": cfg_file,
        "source_sha256": source_digest(asset_root / xml_rel, ik_root / cfg_file),
        "robot": robot.to_dict(),
        "ik_config": cfg.to_dict(),
    }


_PACK_CACHE: Dict[Tuple[str, str], dict] = {}


def load_pack(src_human: str, tgt_robot: str) -> Tuple[RobotModel, IKConfig, dict]:
    """(RobotModel, IKConfig, raw pack dict) for a registered pair.  Unknown keys raise
    KeyError exactly where the reference's dict lookups would (motion_retarget.py:24,30)."""
    _ = ROBOT_XML_REL[tgt_robot]
    _ = IK_CONFIG_REL[src_human][tgt_robot]
    key = (src_human, tgt_robot)
    if key not in _PACK_CACHE:
        p = pack_path(src_human, tgt_robot)
        d = None
        if p.is_file():
            with open(p) as f:
                d = json.load(f)
            d["loaded_from"] = str(p)
            # a checkout next to the pack is the source of truth (ROBOT_XML_DICT / IK_CONFIG_DICT point at ITS files):
            # if its MJCF / IK JSON differ from what the pack was compiled from, compile the checkout's files instead
            root = reference_root()
            if root is not None:
                asset_root, ik_root = _paths(root)
                try:
                    live = source_digest(asset_root / d["xml_rel"], ik_root / d["ik_config_file"])
                except OSError:
                    live = None
                if live is not None and live != d.get("source_sha256"):
                    warnings.warn(f"model pack {p.name} is stale against {root} (source digest differs): compiling the "
                                  f"checkout's MJCF / IK config instead; refresh with tools/compile_packs.py", stacklevel=2)
                    d = None
        if d is None:
            d = compile_pack(src_human, tgt_robot)
            d["loaded_from"] = str(reference_root())
        _PACK_CACHE[key] = d
    d = _PACK_CACHE[key]
    return RobotModel.from_dict(d["robot"]), IKConfig.from_dict(d["ik_config"]), d
