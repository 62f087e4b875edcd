"""stand-in for configargparse: plain argparse"""
from argparse import *  # noqa: F401,F403
from argparse import ArgumentParser as ArgParser  # noqa: F401
