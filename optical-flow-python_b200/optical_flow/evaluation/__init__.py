from optical_flow.evaluation.metrics import flow_angular_error

__all__ = ["flow_angular_error"]
