"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's algorithm for the hot path (plain torch.nn / numpy), used as the parity checker
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under torchok_b200/
may import this package.

Pinning status:
  * retrieval metric (oracle/retrieval.py): PINNED by the reference's own known-answer vectors
    (tests/base_tests/metrics/representation/data.py:123-148,197-230,312-329), reproduced in tests/golden/.
  * JointLoss (oracle/losses.py): PINNED by tests/base_tests/losses/test_base_losses.py:19-77.
  * model path (oracle/models.py): the reference's tests are shape-only and the arithmetic lives in timm 0.6.13 /
    torch, neither importable from /root/reference here => "parity unpinned" by reference goldens; the restatement is
    cross-checked bit-for-bit against torchvision's independent ResNet implementation and against the reference's
    shape contracts (tests/additional_tests/models/backbones/test_backbone.py:140-158).
"""
