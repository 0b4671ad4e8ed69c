"""placeholder -- replaced below"""
__all__ = []
