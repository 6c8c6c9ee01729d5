"""Version of the badread_b200 package (tracks the Badread release whose `simulate` surface it mirrors)."""
__version__ = '0.4.2+b200.1'
