"""NearestNeighbor matcher plugin -- same contract as reference ``localization/matchers/nearest_neighbor.py``
(``descriptors0/1`` as [B, D, N], ``matches0`` / ``matching_scores0`` out); the similarity matrix runs on the
tcgen05 GEMM and top-2 / ratio / distance / mutual check on the device (pram_nn_match)."""
import torch

from ..base_model import BaseModel
from ... import ops


class NearestNeighbor(BaseModel):
    default_conf = {
        'ratio_threshold': None,
        'distance_threshold': None,
        'do_mutual_check': True,
    }
    required_inputs = ['descriptors0', 'descriptors1']

    def _init(self, conf):
        pass

    def _forward(self, data):
        with torch.no_grad():
            m0, s0 = ops.nearest_neighbor_match(data['descriptors0'], data['descriptors1'], self.conf['ratio_threshold'],
                                                self.conf['distance_threshold'], self.conf['do_mutual_check'])
        return {'matches0': m0, 'matching_scores0': s0}
