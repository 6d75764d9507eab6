from .mg_head import Head, MultiGroupHead
