"""pfp-b200: B200-native prefix-free parsing (the newscan.cpp / pscan.cpp stage of Big-BWT).

The directory name carries a hyphen, so load it with `__graft_entry__.load_package()`,
which registers it as the module `bigbwt_b200`.
"""
