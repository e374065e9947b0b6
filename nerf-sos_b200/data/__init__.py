"""Input side of the render path: the reference's on-disk scene format, its dataset classes and
collaters, a device-resident patch sampler, and a synthetic-scene writer for offline runs."""
from .collater import ExhibitCollater, PatchBatchCollater, RayBatchCollater, ViewBatchCollater  # noqa: F401
from .datasets import ExhibitNeRFDataset, PatchNeRFDataset, RayNeRFDataset, ViewNeRFDataset  # noqa: F401
from .synthetic import write_synthetic_scene  # noqa: F401
