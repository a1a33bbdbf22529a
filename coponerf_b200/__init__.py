"""coponerf_b200: sm_100a render path for CoPoNeRF behind the reference's forward()."""
__all__ = ["CoPoNeRF", "RenderEngine"]


def __getattr__(name):
    if name == "CoPoNeRF":
        from .model import CoPoNeRF
        return CoPoNeRF
    if name == "RenderEngine":
        from .render import RenderEngine
        return RenderEngine
    raise AttributeError(name)
