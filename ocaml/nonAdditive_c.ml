(* nonAdditive_c.ml -- device-backed body for lib/nonAdditive_c.ml. The reference implements
   only median_2 (pure OCaml Fitch over int arrays, lib/nonAdditive_c.ml:19-35); here every
   member of NodeData.S (lib/nodeData.ml:3-35) is filled in over the B200 engine's stubs.
   NOT COMPILED IN THIS REPOSITORY (no OCaml toolchain in the build image); the stubs are driven
   from C the same way in tests/c/stub_lifetime.c.

   Node values are [node] custom blocks (slot + engine, finalizer = phylo_fitch_node_release), like
   the reference's Bitvector.t (a custom block over a malloc'd vect, lib/bitvector/bv.c:229-244);
   the spec travels in every [t]: several non-additive character sets = several specs. *)
open Internal

type engine = Likelihood_c.engine
type node = Likelihood_c.node
type codes = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t
type vector = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type ids = Likelihood_c.ids

external set_tips_ : engine -> codes -> int -> vector option -> int -> unit = "nonadd_CAML_set_tips"
external tip_ : engine -> int -> node = "nonadd_CAML_tip"
external median2_ : engine -> node -> node -> node * int = "nonadd_CAML_median2"
external median3_ : engine -> node -> node -> node -> node -> node = "nonadd_CAML_median3"
external distance_ : engine -> node -> node -> int = "nonadd_CAML_distance"
external union_ : engine -> node -> node -> node = "nonadd_CAML_union"
external inter_ : engine -> node -> node -> node = "nonadd_CAML_inter"
external popcount_ : engine -> node -> int = "nonadd_CAML_popcount"
external eltcount_ : engine -> node -> int -> int = "nonadd_CAML_eltcount"
external compare_ : engine -> node -> node -> int = "nonadd_CAML_compare"
external score_tree_ : engine -> ids -> int -> int -> int * node array = "nonadd_CAML_score_tree"
external uppass_ : engine -> ids -> int -> int -> unit = "nonadd_CAML_uppass"
external get_states_ :
  engine -> node -> bool -> (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array1.t -> unit
  = "nonadd_CAML_get_states"

type m = unit

type spec = { engine : engine; n_taxa : int; n_chars : int; }

(* [cost] is node-local, like the reference's (lib/nonAdditive_c.ml:35, lib/node.ml:191) *)
type t = { spec : spec; node : node; codes : IntSet.t; cost : float; }

let filter_codes set t =
  let c = IntSet.inter set t.codes in if IntSet.is_empty c then None else Some { t with codes = c }
let filter_codes_comp set t =
  let c = IntSet.diff t.codes set in if IntSet.is_empty c then None else Some { t with codes = c }
let cardinal t = t.spec.n_chars
let get_codes t = t.codes
let mem codes t = match codes with
  | None -> true | Some xs -> List.exists (fun x -> IntSet.mem x t.codes) xs
(* union of the children's state sets (bv_union, lib/bitvector/bv.c:93-99), a new node value *)
let union _p a b =
  { a with node = union_ a.spec.engine a.node b.node; cost = 0.0; codes = IntSet.union a.codes b.codes }
let compare a b = compare_ a.spec.engine a.node b.node
let recode f t = { t with codes = IntSet.fold (fun x acc -> IntSet.add (f x) acc) t.codes IntSet.empty }

let median_1 _ _ x = x                                   (* lib/nonAdditive_c.ml:18 *)
let median_2 _ _ x y =                                   (* lib/nonAdditive_c.ml:19-35 *)
  let n, c = median2_ x.spec.engine x.node y.node in
  { x with node = n; cost = float_of_int c; codes = IntSet.union x.codes y.codes }
(* final states of a node (Fitch's second pass): [prev] = the node's own preliminary value, a = its
   parent's final value, b and c = its children's preliminary values. Without [prev] the preliminary
   value is computed first. The reference only sketches this (bv_CAML_fitch_median3, bv.h:94). *)
let median_3 m prev a b c =
  let own = match prev with Some o -> o | None -> median_2 m None b c in
  { own with node = median3_ own.spec.engine own.node a.node b.node c.node; cost = own.cost }
let median_n m prev x = function
  | [y] -> median_2 m prev x y
  | [y; z] -> median_3 m prev x y z
  | _ -> failwith "NonAdditive_c.median_n: binary trees only"
(* the final-state pass never changes a parsimony length: the node is its own best adjustment; the set of
   characters whose final sets differ from the preliminary ones is reported as changed *)
let adjust_3 m _codes n a b c =
  let f = median_3 m (Some n) a b c in
  if compare f n = 0 then n, IntSet.empty else f, n.codes
let adjust_n m codes n = function
  | [a; b; c] -> adjust_3 m codes n a b c
  | _ -> n, IntSet.empty
let cost t = t.cost
let root_cost t = t.cost
let leaf_cost _ = 0.0
let distance_1 _ a b = float_of_int (distance_ a.spec.engine a.node b.node)   (* bv_distance *)
(* a joined to the edge (b, c): bv_distance to the median of the edge's two ends *)
let distance_2 m a b c = distance_1 m a (median_2 m None b c)
let to_string t =
  let n = t.spec.n_chars in
  let out = Bigarray.Array1.create Bigarray.int8_unsigned Bigarray.c_layout n in
  get_states_ t.spec.engine t.node false out;
  String.concat "," (List.init n (fun i -> string_of_int out.{i}))

let of_string _ = failwith "NonAdditive_c.of_string: use create_spec"
let of_parser _ = failwith "NonAdditive_c.of_parser: use create_spec"

(* one engine per character set *)
let create_spec ?(device = 0) (chars : codes) n_states (weights : vector option) =
  let e = Likelihood_c.engine_create device in
  let n_taxa = Bigarray.Array2.dim1 chars in
  set_tips_ e chars n_states weights (2 * n_taxa);
  { engine = e; n_taxa; n_chars = Bigarray.Array2.dim2 chars }

let leaf (s : spec) (i : int) (code : int) =
  { spec = s; node = tip_ s.engine i; codes = IntSet.singleton code; cost = 0.0 }

(* whole tree in one launch: (length, node values in schedule order); [uppass] then leaves final sets
   readable with get_states_ ... true *)
let score_tree (s : spec) (ops : ids) root_a root_b = score_tree_ s.engine ops root_a root_b
(* final sets for the tree [score_tree] just scored: the interior ids of [ops] are rewritten to the slots
   the returned node values hold (schedule order), which is how the engine names them *)
let uppass (s : spec) (ops : ids) (nodes : node array) root_a root_b =
  let n = Bigarray.Array2.dim1 ops in
  let slot_of = Hashtbl.create n in
  for i = 0 to n - 1 do
    Hashtbl.replace slot_of (Int32.to_int ops.{i, 0}) (Likelihood_c.node_slot nodes.(i))
  done;
  let map id = try Hashtbl.find slot_of id with Not_found -> id in
  let real = Bigarray.Array2.create Bigarray.int32 Bigarray.c_layout n 3 in
  for i = 0 to n - 1 do
    for j = 0 to 2 do real.{i, j} <- Int32.of_int (map (Int32.to_int ops.{i, j})) done
  done;
  uppass_ s.engine real (map root_a) (map root_b)
let popcount t = popcount_ t.spec.engine t.node            (* bv_popcount, bv.c:102-108 *)
let eltcount t i = eltcount_ t.spec.engine t.node i        (* bv_eltcount, bv.c:59-69 *)
let inter a b = { a with node = inter_ a.spec.engine a.node b.node; cost = 0.0 }
