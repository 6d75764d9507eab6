"""Box coders needed at config-load / predict time.

GroundBox3dCoderTorch mirrors det3d/core/bbox/box_coders.py:32-44,100-109:
`.code_size`, `.n_dim`, `.decode_torch(encodings, anchors)`.  Encoding (training
target assignment) is out of scope for the inference hot path.
"""
from . import box_torch_ops


class GroundBox3dCoderTorch:
    def __init__(self, linear_dim=False, vec_encode=False, n_dim=7, norm_velo=False):
        self.linear_dim = linear_dim
        self.vec_encode = vec_encode
        self.norm_velo = norm_velo
        self.n_dim = n_dim

    @property
    def code_size(self):
        return self.n_dim + 1 if self.vec_encode else self.n_dim

    def decode_torch(self, boxes, anchors):
        # Reference quirk, reproduced on purpose (box_coders.py:106-109): `linear_dim` is passed POSITIONALLY and
        # lands in second_box_decode's ignored `bin_loss` slot (box_torch_ops.py:80-87), and `norm_velo` is never
        # forwarded -- so the reference always decodes sizes with exp() and velocities without the diagonal,
        # whatever the coder was built with.  The fused kernel (mg_head.predict_device) follows the same rule.
        return box_torch_ops.second_box_decode(boxes, anchors, self.vec_encode, self.linear_dim)

    def encode_torch(self, boxes, anchors):
        raise NotImplementedError("box encoding is training-only; det3d_b200 covers the inference path")
