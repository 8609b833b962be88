def __getattr__(name):
    raise NotImplementedError(f"matplotlib.pyplot.{name} is not available in the test shim")
