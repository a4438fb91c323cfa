# placeholder, replaced below
def sigma0_detrend(*a, **k):
    raise NotImplementedError
