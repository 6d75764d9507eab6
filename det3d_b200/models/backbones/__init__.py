from .scn import SparseBasicBlock, SpMiddleFHD, SpMiddleResNetFHD
