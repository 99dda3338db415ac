def rdchiralRunText(*_a, **_k):
    raise RuntimeError("rdchiral stand-in: template application is outside the oracle's scope")
