def __getattr__(name):
    raise NotImplementedError("matplotlib is not available (oracle shim)")
