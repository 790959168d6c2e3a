from .cohort import DeviceCohort
from .loader import AsyncBagLoader, PackedBatch, pack_bags
from .store import PatchFeatureStore, WSIPatchSurvStore, build_store, build_store_from_files, read_patch_data

__all__ = ["AsyncBagLoader", "DeviceCohort", "PackedBatch", "pack_bags", "PatchFeatureStore", "WSIPatchSurvStore", "build_store",
           "build_store_from_files", "read_patch_data"]
