"""stand-in for matplotlib (only imported, never drawn with, on the train / test path)"""


def use(*args, **kwargs):
    return None
