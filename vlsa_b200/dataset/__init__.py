from .loader import AsyncBagLoader, PackedBatch, pack_bags

__all__ = ["AsyncBagLoader", "PackedBatch", "pack_bags"]
