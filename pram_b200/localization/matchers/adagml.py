"""Matcher plugin -- same contract as reference ``localization/matchers/adagml.py``: build the net from
``conf``, load ``conf['weight_path']`` (checkpoint key 'model', strict), run under no_grad."""
import torch

from ..base_model import BaseModel
from ...nets.adagml import AdaGML as _Net


class AdaGML(BaseModel):
    default_conf = {}
    required_inputs = [
        'image0', 'keypoints0', 'scores0', 'descriptors0',
        'image1', 'keypoints1', 'scores1', 'descriptors1',
    ]

    def _init(self, conf):
        self.net = _Net(config={k: v for k, v in conf.items() if k not in ('name', 'weight_path')}).eval()
        if conf.get('weight_path'):
            # the shipped checkpoint pickles a numpy scalar ('min_loss'): needs weights_only=False on torch>=2.6
            state = torch.load(conf['weight_path'], map_location='cpu', weights_only=False)['model']
            self.net.load_state_dict(state, strict=True)

    def _forward(self, data):
        with torch.no_grad():
            return self.net(data)
