"""Host-side mirror of the reference's `Matching` front-end (models/matching.py:8-30)."""
import torch

from .frontend import host_stages, sift_forward
from .gmatcher import GMatcher


class Matching(torch.nn.Module):
    """Image Matching Frontend — same constructor and dict-in/dict-out call as the reference.

    With `keypoints{0,1}` in `data` the matcher runs on them directly (matching.py:17,21); without, the images go
    through `sift_forward` first (matching.py:18-24 -> utils/common.py:837-893; here gims_b200/frontend.py: OpenCV SIFT
    and patches on the host, the caller's CAR-HyNet on the device, descriptors never leave it).
    """

    def __init__(self, config={}):
        super().__init__()
        self.gmodel = GMatcher(config)
        self.max_keypoints = config.get('max_keypoints', -1)

    def forward(self, data):
        pred = {}
        need = [s for s in ('0', '1') if 'keypoints' + s not in data]
        # the OpenCV stages of both images run side by side (the reference does image 0, then image 1)
        staged = dict(zip(need, host_stages([data['image' + s] for s in need], self.max_keypoints))) if need else {}
        for s in need:
            dev = data.get('device', self.gmodel.bin_score.device)
            out = sift_forward({'image': data['image' + s], 'max_keypoints': self.max_keypoints,
                                'carhynet': data['carhynet']}, device=dev, staged=staged[s])
            pred = {**pred, **{k + s: v for k, v in out.items()}}
        data = {**data, **pred}
        for k in data:
            if isinstance(data[k], (list, tuple)):
                data[k] = torch.stack(data[k])
        pred = {**pred, **self.gmodel(data)}
        return pred
