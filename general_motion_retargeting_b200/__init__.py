"""B200-native batched motion retargeting (the hot path of GMR)."""
from .params import (IK_CONFIG_DICT, ROBOT_BASE_DICT, ROBOT_XML_DICT,  # noqa: F401
                     VIEWER_CAM_DISTANCE_DICT)

__all__ = ["GeneralMotionRetargeting", "RetargetFailure", "retarget_mixed", "ROBOT_XML_DICT", "IK_CONFIG_DICT",
           "ROBOT_BASE_DICT", "VIEWER_CAM_DISTANCE_DICT"]


def __getattr__(name):
    # GeneralMotionRetargeting pulls in torch + the CUDA extension; import it lazily so
    # that the model compiler and registries stay usable without either.
    if name == "GeneralMotionRetargeting":
        from .motion_retarget import GeneralMotionRetargeting
        return GeneralMotionRetargeting
    if name == "RetargetFailure":
        from .motion_retarget import RetargetFailure
        return RetargetFailure
    if name == "retarget_mixed":
        from .motion_retarget import retarget_mixed
        return retarget_mixed
    raise AttributeError(name)
