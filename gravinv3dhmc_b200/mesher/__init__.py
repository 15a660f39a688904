from .geometry import Prism, Tesseroid  # noqa: F401
from .mesh import PrismMesh, TesseroidMesh, PrismMeshSegment, TesseroidMeshSegment  # noqa: F401
