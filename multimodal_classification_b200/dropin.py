"""Bind the B200 classes under the reference's own names.  Call ``install()`` once, before the
reference's Kedro nodes run; no reference file is edited.

The reference resolves the model inside ``_load_facebook_model`` (pipelines/model_training/nodes.py:223-230) from
``multimodalclassification.models`` and the extractor through the dict literal in
``models/feature_extractors/__init__.py:101-112`` / ``FEATURE_EXTRACTOR_REGISTRY`` (models/base.py:274), and the feature-store
loaders inside ``create_dataloaders_lmdb / create_dataloaders_precomputed`` (nodes.py:605-658) from
``pipelines.data_processing.lmdb_dataset / precomputed_dataset``."""
from __future__ import annotations


def install(ingest: bool = True) -> None:
    import multimodalclassification.models as M
    import multimodalclassification.models.base as B
    import multimodalclassification.models.feature_extractors as FE
    import multimodalclassification.models.vilbert_facebook_arch as A

    from .resnet152_roi import ResNet152ROIExtractor
    from .resnet_grid import ResNetFeatureExtractor, ResNetVGExtractor
    from .fasterrcnn_vg import FasterRCNNVGExtractor
    from .fasterrcnn_vg_rpn import FasterRCNNVGRPNExtractor
    from .vilbert import ViLBERTForClassification, get_facebook_vilbert_config, load_facebook_weights
    from . import vilbert_core as core

    M.ViLBERTFacebookArch = A.ViLBERTForClassification = ViLBERTForClassification
    M.get_facebook_vilbert_config = A.get_facebook_vilbert_config = get_facebook_vilbert_config
    M.load_facebook_weights = A.load_facebook_weights = load_facebook_weights
    # the second two-stream surface (models/vilbert_core.py:593-657; `ViLBERTCore` in models/__init__.py): same engine
    import multimodalclassification.models.vilbert_core as VC
    VC.ViLBERTForClassification = core.ViLBERTForClassification
    VC.get_vilbert_config = core.get_vilbert_config
    for name in ("ViLBERTCore", "ViLBERTCoreForClassification"):
        if hasattr(M, name):
            setattr(M, name, core.ViLBERTForClassification)
    FE.ResNet152ROIExtractor = ResNet152ROIExtractor
    B.FEATURE_EXTRACTOR_REGISTRY["resnet152_roi"] = ResNet152ROIExtractor
    FE.ResNetFeatureExtractor = ResNetFeatureExtractor                 # "resnet" in the same dict literal / registry
    B.FEATURE_EXTRACTOR_REGISTRY["resnet"] = ResNetFeatureExtractor
    FE.ResNetVGExtractor = ResNetVGExtractor                           # "resnet_vg"
    B.FEATURE_EXTRACTOR_REGISTRY["resnet_vg"] = ResNetVGExtractor
    FE.FasterRCNNVGExtractor = FasterRCNNVGExtractor                   # "fasterrcnn_vg" (fasterrcnn_vg.py:170)
    B.FEATURE_EXTRACTOR_REGISTRY["fasterrcnn_vg"] = FasterRCNNVGExtractor
    FE.FasterRCNNVGRPNExtractor = FasterRCNNVGRPNExtractor             # "fasterrcnn_vg_rpn" (fasterrcnn_vg_rpn.py:290)
    B.FEATURE_EXTRACTOR_REGISTRY["fasterrcnn_vg_rpn"] = FasterRCNNVGRPNExtractor
    if ingest:
        from . import ingest as I
        for module, name in (("lmdb_dataset", "create_lmdb_dataloaders"), ("precomputed_dataset", "create_precomputed_dataloaders")):
            try:        # these modules import lmdb / h5py / kedro; where those are absent the pipelines cannot run at all
                mod = __import__("multimodalclassification.pipelines.data_processing." + module, fromlist=[name])
            except ImportError:
                continue
            setattr(mod, name, getattr(I, name))
