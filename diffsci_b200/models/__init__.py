# flake8: noqa
"""Mirror of diffsci.models for the Karras/EDM hot path."""
from .karras import *
from .nets import *
