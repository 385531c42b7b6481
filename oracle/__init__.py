"""CPU oracle for the Complex-YOLOv4 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker.  The product (complex-yolov4-pytorch_b200/)
never imports it and fails loudly when its CUDA library is missing.
"""
