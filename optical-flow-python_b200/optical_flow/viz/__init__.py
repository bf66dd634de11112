"""Flow visualisation edge kept on the device: Middlebury colour coding (reference: viz/flow_color.py)."""
from optical_flow.viz.flow_color import flow_to_color

__all__ = ['flow_to_color']
