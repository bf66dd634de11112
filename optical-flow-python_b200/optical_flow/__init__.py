"""optical_flow -- B200-native drop-in for the coarse-to-fine HS / BA / Classic+NL hot path of
jordanshivers/optical-flow-python.  Same public names as the reference package for that path; every numerical
stage runs as hand-written sm_100a CUDA behind libb200flow.so (include/b200flow.h).  No CPU fallback."""
from optical_flow.interface import estimate_flow, estimate_flow_batch, estimate_flow_sharded
from optical_flow.io.flo_io import read_flo, write_flo, flo_bytes
from optical_flow.evaluation.metrics import flow_angular_error, flow_error_batch
from optical_flow.viz.flow_color import flow_to_color
from optical_flow.methods.config import load_of_method

__all__ = ['estimate_flow', 'estimate_flow_batch', 'estimate_flow_sharded', 'read_flo', 'write_flo', 'flo_bytes',
           'flow_angular_error', 'flow_error_batch', 'flow_to_color', 'load_of_method']
