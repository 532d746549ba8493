"""Stand-in for pyspoa so reference `local_asm.py` imports; POA itself is not available here."""


def poa(*a, **k):
    raise RuntimeError("stub spoa: pyspoa is not installed in this container")
