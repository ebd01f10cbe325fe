"""The drop-in binding against the real reference package (authoring container only: /root/reference does not travel)."""
import os
import sys

import pytest

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the authoring container")


def test_install_rebinds_the_reference_lookups():
    sys.path.insert(0, REF)
    import torch
    import multimodalclassification.models as M
    from multimodalclassification.models.vilbert_facebook_arch import ViLBERTForClassification as RefModel
    from multimodalclassification.models.feature_extractors import get_feature_extractor
    from multimodal_classification_b200 import dropin, vilbert
    from multimodal_classification_b200._lib import VbError
    from oracle import vilbert_oracle as vo

    cfg = vo.tiny_config()
    torch.manual_seed(0)
    ref = RefModel(cfg, num_labels=2)
    dropin.install()
    assert M.ViLBERTFacebookArch is vilbert.ViLBERTForClassification
    # what nodes.py:223-230 does
    from multimodalclassification.models import ViLBERTFacebookArch, get_facebook_vilbert_config
    assert get_facebook_vilbert_config() == vilbert.get_facebook_vilbert_config()
    ours = ViLBERTFacebookArch(cfg, num_labels=2)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())
    assert all(sd_ref[k].shape == sd_ours[k].shape and sd_ref[k].dtype == sd_ours[k].dtype for k in sd_ref)
    assert ours.load_state_dict(sd_ref, strict=True) is not None
    assert ours.get_num_parameters() == ref.get_num_parameters()
    ours.freeze_bert_layers(2); ref.freeze_bert_layers(2)
    assert [p.requires_grad for p in ours.parameters()] == [p.requires_grad for p in ref.parameters()]
    # the extractor factory now reaches our class (which refuses a CPU device instead of silently falling back)
    with pytest.raises(VbError):
        get_feature_extractor("resnet152_roi", device="cpu")
    for name in ("resnet", "resnet_vg", "fasterrcnn_vg", "fasterrcnn_vg_rpn"):
        with pytest.raises(VbError):
            get_feature_extractor(name, device="cpu")
