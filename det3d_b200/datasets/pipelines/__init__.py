from .compose import Compose
from .formating import Reformat
from .loading import LoadPointCloudAnnotations, LoadPointCloudFromFile, ingest_sweeps, read_file
from .preprocess import AssignTarget, Preprocess, Voxelization
