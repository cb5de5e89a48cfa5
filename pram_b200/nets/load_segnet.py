"""Drop-in for reference ``nets/load_segnet.py:12-31`` (only the ViT recogniser is on the hot path)."""
from .segnetvit import SegNetViT


def load_segnet(network, n_class, desc_dim, n_layers, output_dim):
    cfg = {'descriptor_dim': desc_dim, 'n_layers': n_layers, 'n_class': n_class, 'output_dim': output_dim,
           'with_score': False}
    if network == 'segnetvit':
        return SegNetViT(cfg)
    raise ValueError(f'{network}: only "segnetvit" is implemented (the Conv1d "segnet" recogniser is superseded '
                     f'in every shipped config, SURVEY.md section 2)')
