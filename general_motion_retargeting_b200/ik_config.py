"""IK config JSON → flat task tables.

Mirrors what the reference constructor derives from ``ik_configs/*.json``
(reference general_motion_retargeting/motion_retarget.py:30-114):

* ``human_scale_table`` entries are multiplied by ``actual_human_height /
  human_height_assumption`` (:36-43) — kept separate here (`scale` × per-clip `ratio`)
  so one compiled table serves a batch of clips with different heights;
* one ``mink.FrameTask(frame_name, "body", position_cost, orientation_cost,
  lm_damping=1)`` per table entry with a non-zero weight (:80-89, :98-107);
* ``pos_offsets1[body] = pos_offset - ground_height*z`` and ``rot_offsets1[body] =
  R.from_quat(rot_offset, scalar_first=True)`` (normalising) (:91-94).  Table-2
  offsets are built by the reference (:109-112) but never applied (:121) — dropped.

Human bodies are indexed in ``human_scale_table`` key order: that is the ``nh`` axis of
the batched ``pos[C,T,nh,3]`` / ``quat[C,T,nh,4]`` inputs.
"""
from __future__ import annotations

import copy
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .mjcf import RobotModel


@dataclass
class IKConfig:
    """The reference's ik_config dict, parsed (same keys as the JSON)."""

    robot_root_name: str
    human_root_name: str
    ground_height: float
    human_height_assumption: float
    use_ik_match_table1: bool
    use_ik_match_table2: bool
    human_scale_table: Dict[str, float]
    ik_match_table1: Dict[str, list]
    ik_match_table2: Dict[str, list]

    @staticmethod
    def from_dict(d: dict) -> "IKConfig":
        return IKConfig(
            robot_root_name=d["robot_root_name"],
            human_root_name=d["human_root_name"],
            ground_height=float(d["ground_height"]),
            human_height_assumption=float(d["human_height_assumption"]),
            use_ik_match_table1=bool(d["use_ik_match_table1"]),
            use_ik_match_table2=bool(d["use_ik_match_table2"]),
            human_scale_table={k: float(v) for k, v in d["human_scale_table"].items()},
            ik_match_table1=copy.deepcopy(d["ik_match_table1"]),
            ik_match_table2=copy.deepcopy(d["ik_match_table2"]),
        )

    @staticmethod
    def from_json(path: str) -> "IKConfig":
        with open(os.fspath(path)) as f:
            return IKConfig.from_dict(json.load(f))

    def to_dict(self) -> dict:
        return {
            "robot_root_name": self.robot_root_name,
            "human_root_name": self.human_root_name,
            "ground_height": self.ground_height,
            "human_height_assumption": self.human_height_assumption,
            "use_ik_match_table1": self.use_ik_match_table1,
            "use_ik_match_table2": self.use_ik_match_table2,
            "human_scale_table": dict(self.human_scale_table),
            "ik_match_table1": copy.deepcopy(self.ik_match_table1),
            "ik_match_table2": copy.deepcopy(self.ik_match_table2),
        }


@dataclass
class TaskTable:
    """Device-ready constants for one (source format, robot) pair."""

    human_names: List[str]      # [nh] scale-table order
    root_idx: int               # index of human_root_name in human_names
    scale: np.ndarray           # [nh] float64, before the per-clip height ratio
    height_assumption: float
    pos_off: np.ndarray         # [nh,3] table-1 local offset minus ground_height*z
    rot_off: np.ndarray         # [nh,4] wxyz, normalised
    foot_mask: np.ndarray       # [nh] bool: name contains "Foot"/"foot" (offset_to_ground, :252-270)
    task_frames: List[str]      # [nt] robot body names
    task_body: np.ndarray       # [nt] int32 robot body index
    task_human: np.ndarray      # [nt] int32 human body index
    w1: np.ndarray              # [nt,2] (position_cost, orientation_cost) in stage 1, 0 if absent
    w2: np.ndarray              # [nt,2] same for stage 2
    in1: np.ndarray             # [nt] bool: task belongs to tasks1
    in2: np.ndarray             # [nt] bool: task belongs to tasks2
    use1: bool
    use2: bool

    @property
    def nh(self) -> int:
        return len(self.human_names)

    @property
    def nt(self) -> int:
        return len(self.task_frames)


def _active_entries(table: Dict[str, list]) -> List[Tuple[str, str, float, float, list, list]]:
    out = []
    for frame_name, entry in table.items():
        body_name, pos_w, rot_w, pos_off, rot_off = entry
        if pos_w != 0 or rot_w != 0:     # motion_retarget.py:82, :100
            out.append((frame_name, body_name, float(pos_w), float(rot_w), pos_off, rot_off))
    return out


def compile_task_table(robot: RobotModel, cfg: IKConfig) -> TaskTable:
    human_names = list(cfg.human_scale_table.keys())
    if cfg.human_root_name not in cfg.human_scale_table:
        # scale_human_data indexes human_scale_table[human_root_name] (:215)
        raise KeyError(cfg.human_root_name)
    hidx = {n: i for i, n in enumerate(human_names)}
    nh = len(human_names)

    e1 = _active_entries(cfg.ik_match_table1)
    e2 = _active_entries(cfg.ik_match_table2)

    # The reference keys tasks and offsets by HUMAN body name (:90-94); two robot frames
    # bound to one human body would leave the first task without a target (mink raises
    # TargetNotSet).  Refuse such a table up front.
    for tab, ent in (("ik_match_table1", e1), ("ik_match_table2", e2)):
        names = [e[1] for e in ent]
        if len(names) != len(set(names)):
            raise ValueError(f"{tab}: a human body is bound to more than one robot frame")

    pos_off = np.zeros((nh, 3))
    rot_off = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (nh, 1))
    have = np.zeros(nh, bool)
    ground = cfg.ground_height * np.array([0.0, 0.0, 1.0])
    for (_, body, _, _, p_off, r_off) in e1:
        if body in hidx:
            i = hidx[body]
            pos_off[i] = np.asarray(p_off, np.float64) - ground
            q = np.asarray(r_off, np.float64)
            rot_off[i] = q / np.linalg.norm(q)       # scipy R.from_quat normalises
            have[i] = True
    # offset_human_data looks up pos_offsets1[body] for every scaled body (:237-244)
    for i, n in enumerate(human_names):
        if not have[i]:
            raise KeyError(n)

    tasks: Dict[Tuple[str, str], int] = {}
    frames: List[str] = []
    tb: List[int] = []
    th: List[int] = []
    w1: List[List[float]] = []
    w2: List[List[float]] = []
    in1: List[bool] = []
    in2: List[bool] = []

    def slot(frame: str, body: str) -> int:
        key = (frame, body)
        if key not in tasks:
            if body not in hidx:
                # update_targets reads human_data[body] after scale_human_data dropped it (:129,:135)
                raise KeyError(body)
            tasks[key] = len(frames)
            frames.append(frame)
            tb.append(robot.body_id(frame))
            th.append(hidx[body])
            w1.append([0.0, 0.0]); w2.append([0.0, 0.0]); in1.append(False); in2.append(False)
        return tasks[key]

    if cfg.use_ik_match_table1:
        for (frame, body, pw, rw, _, _) in e1:
            k = slot(frame, body); w1[k] = [pw, rw]; in1[k] = True
    if cfg.use_ik_match_table2:
        for (frame, body, pw, rw, _, _) in e2:
            k = slot(frame, body); w2[k] = [pw, rw]; in2[k] = True

    return TaskTable(
        human_names=human_names,
        root_idx=hidx[cfg.human_root_name],
        scale=np.array([cfg.human_scale_table[n] for n in human_names], np.float64),
        height_assumption=cfg.human_height_assumption,
        pos_off=pos_off,
        rot_off=rot_off,
        foot_mask=np.array([("Foot" in n) or ("foot" in n) for n in human_names], bool),
        task_frames=frames,
        task_body=np.asarray(tb, np.int32),
        task_human=np.asarray(th, np.int32),
        w1=np.asarray(w1, np.float64).reshape(-1, 2),
        w2=np.asarray(w2, np.float64).reshape(-1, 2),
        in1=np.asarray(in1, bool),
        in2=np.asarray(in2, bool),
        use1=cfg.use_ik_match_table1,
        use2=cfg.use_ik_match_table2,
    )
