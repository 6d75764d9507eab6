from .loading import LoadPointCloudFromFile, ingest_sweeps, read_file
