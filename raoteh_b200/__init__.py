"""raoteh_b200: B200-native hot path of argriffing/raoteh (tree MJP likelihood,
closed-form expectations, Rao-Teh sweeps), batched over sites and chains."""
__version__ = '0.1.0'
