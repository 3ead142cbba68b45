class FactorizedTensor:
    def __init__(self):
        pass
