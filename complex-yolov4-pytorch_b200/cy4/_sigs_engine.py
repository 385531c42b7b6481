"""ctypes signatures of the conv-engine entry points (filled in as csrc/ grows)."""
SIGS = {}
