"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithm for the hot path (plain torch.nn / numpy), used as the parity checker
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under torchok_b200/
may import this package.

Pinning status:
  * retrieval metric (oracle/retrieval.py): PINNED by the reference's own known-answer vectors
    (tests/base_tests/metrics/representation/data.py:123-148,197-230,312-329), reproduced in tests/golden/.
  * JointLoss (oracle/losses.py): PINNED by tests/base_tests/losses/test_base_losses.py:19-77.
  * units, heads, necks and losses that the reference implements with torch alone (oracle/models.py: ConvBnAct,
    LinearHead, ClassificationHead, ArcFaceHead, SegmentationHead, HRNetSegmentationNeck, ContrastiveLoss,
    calc_relevance_matrix, DiceLoss): PINNED against outputs of the reference's own files, executed by path under
    stubbed framework imports (tests/golden/make_reference_goldens.py -> tests/golden/reference_goldens.pt).
  * networks whose arithmetic lives in timm 0.6.13 / mmdet 3.0.0 (ResNet, Swin-V2, HRNet, FPN, poolings): the
    reference's tests are shape-only and those packages are not importable here => "parity unpinned" by reference
    goldens.  Cross-checks against independent implementations in this image: ResNet-18/50 bit-for-bit against
    torchvision, the Swin-V2 block, PatchMerging and the WHOLE network against torchvision's SwinTransformer V2, FPN
    against torchvision.ops.FeaturePyramidNetwork (tests/test_oracle_models.py); HRNet has no second implementation
    here and is held only to the reference's shape contracts (tests/additional_tests/models/backbones/test_backbone.py).
"""
