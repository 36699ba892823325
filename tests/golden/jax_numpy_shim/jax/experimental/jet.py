def jet(*_a, **_k):
    raise NotImplementedError("Taylor-mode autodiff (odetools/init.py) is not on the EK1 path and not in the shim")
