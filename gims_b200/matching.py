"""Host-side mirror of the reference's `Matching` front-end (models/matching.py:8-30)."""
import torch

from .gmatcher import GMatcher


class Matching(torch.nn.Module):
    """Image Matching Frontend — same constructor and dict-in/dict-out call as the reference.

    The SIFT + CAR-HyNet front-end (`sift_forward`, utils/common.py:837-893) that the reference runs
    when `keypoints{0,1}` are absent is outside this build's scope (SURVEY.md §8f "next" row 1):
    callers supply keypoints/descriptors/scores, which is the path matching.py:17,21 takes as well.
    """

    def __init__(self, config={}):
        super().__init__()
        self.gmodel = GMatcher(config)
        self.max_keypoints = config.get('max_keypoints', -1)

    def forward(self, data):
        pred = {}
        if 'keypoints0' not in data or 'keypoints1' not in data:
            raise NotImplementedError('feature extraction (sift_forward + CAR-HyNet) is not part of the B200 hot '
                                      'path; pass keypoints*/descriptors*/scores* as matching.py:17,21 allow')
        data = {**data, **pred}
        for k in data:
            if isinstance(data[k], (list, tuple)):
                data[k] = torch.stack(data[k])
        pred = {**pred, **self.gmodel(data)}
        return pred
