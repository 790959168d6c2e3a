from .vlsa_handler import VLSAHandler, create_output_converter, fetch_kws

__all__ = ["VLSAHandler", "create_output_converter", "fetch_kws"]
