import torch


def _assert(condition, message):
    torch._assert(condition, message)
