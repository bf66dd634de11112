from optical_flow.io.flo_io import read_flo, write_flo, read_flow_file

__all__ = ["read_flo", "write_flo", "read_flow_file"]
