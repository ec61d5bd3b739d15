import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from hnumo_loader import hnumo_b200 as hn
p = dict(hn.decks.SHIPPED["double_gyre"], nelx=4, nely=4)
deck = hn.decks.build_deck(p)
S = hn.Solver(deck)
S.set_option("use_graph", 0)
S.upload_state(deck["q_df"], deck["qb_df"], deck["qprime_df"])
print("step rc", S.step(1))
