"""torch_de_b200: the PINN residual hot path of torch_DE (jets forward, residual/loss, reverse to the
weight gradient) as hand-written sm_100a CUDA behind torch_DE's own engine / loss API."""
from .data import PINN_dict, PINN_group, make_batch, PINN_sampler, DeviceLoader
from .de_getter import DE_Getter
from .engine import FD_engine, Jet_engine, JetRunner
from .equations import DE_func, get_derivatives, get_NavierStokes
from .loss import Loss, Loss_handler
from .grad_balance import GradNorm, Causal_weighting
from .graph_step import GraphedStep
from .point_weights import RegularGridInterpolator, R3_sampler, sample_from_tensor
from .networks import MLP, Fourier_Embedding, Fourier_Net, RWF_Linear, Sin, extract_layers
from .spec import JetChannels, derivative_index_map

__all__ = ['DE_Getter', 'Jet_engine', 'FD_engine', 'GradNorm', 'Causal_weighting', 'RegularGridInterpolator', 'R3_sampler', 'sample_from_tensor', 'PINN_sampler', 'DeviceLoader', 'GraphedStep', 'JetRunner', 'JetChannels', 'derivative_index_map', 'MLP', 'Fourier_Net', 'Fourier_Embedding', 'RWF_Linear', 'Sin', 'extract_layers',
           'DE_func', 'get_derivatives', 'get_NavierStokes', 'Loss', 'Loss_handler', 'PINN_dict', 'PINN_group', 'make_batch']
