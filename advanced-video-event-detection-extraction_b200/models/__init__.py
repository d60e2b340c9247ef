from .openclip_model import OpenCLIPModel  # noqa: F401
