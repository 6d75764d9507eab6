from .rpn import RPN
