(* nonAdditive_c.ml -- device-backed body for lib/nonAdditive_c.ml. The reference implements
   only median_2 (pure OCaml Fitch over int arrays, lib/nonAdditive_c.ml:19-35); here every
   member of NodeData.S (lib/nodeData.ml:3-35) is filled in over the B200 engine's stubs.
   NOT COMPILED IN THIS REPOSITORY (no OCaml toolchain in the build image). *)
open Internal

type engine = Likelihood_c.engine
type codes = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t
type vector = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t

external set_tips_ : engine -> codes -> int -> vector option -> int -> unit = "nonadd_CAML_set_tips"
external median2_ : engine -> int -> int -> int -> int = "nonadd_CAML_median2"
external distance_ : engine -> int -> int -> int = "nonadd_CAML_distance"
external union_ : engine -> int -> int -> int -> unit = "nonadd_CAML_union"
external compare_ : engine -> int -> int -> int = "nonadd_CAML_compare"
external get_states_ :
  engine -> int -> bool -> (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array1.t -> unit
  = "nonadd_CAML_get_states"

type m = unit

(* [cost] is node-local, like the reference's (lib/nonAdditive_c.ml:35, lib/node.ml:191) *)
type t = { eng : engine; slot : int; codes : IntSet.t; cost : float; n_chars : int; }

type spec = { engine : engine; capacity : int; mutable next : int; }
let the_spec : spec option ref = ref None
let fresh () = match !the_spec with
  | None -> failwith "NonAdditive_c: create_spec first"
  | Some s -> let i = s.next in
    if i >= s.capacity then failwith "NonAdditive_c: node capacity exhausted";
    s.next <- i + 1; i

let filter_codes set t =
  let c = IntSet.inter set t.codes in if IntSet.is_empty c then None else Some { t with codes = c }
let filter_codes_comp set t =
  let c = IntSet.diff t.codes set in if IntSet.is_empty c then None else Some { t with codes = c }
let cardinal t = t.n_chars
let get_codes t = t.codes
let mem codes t = match codes with
  | None -> true | Some xs -> List.exists (fun x -> IntSet.mem x t.codes) xs
let union _ a b = let p = fresh () in union_ a.eng p a.slot b.slot; { a with slot = p; cost = 0.0 }
let compare a b = compare_ a.eng a.slot b.slot
let recode f t = { t with codes = IntSet.fold (fun x acc -> IntSet.add (f x) acc) t.codes IntSet.empty }

let median_1 _ _ x = x                                   (* lib/nonAdditive_c.ml:18 *)
let median_2 _ _ x y =                                   (* lib/nonAdditive_c.ml:19-35 *)
  let p = fresh () in
  let c = median2_ x.eng p x.slot y.slot in
  { x with slot = p; cost = float_of_int c; codes = IntSet.union x.codes y.codes }
let median_3 m prev _ x y = median_2 m prev x y
let median_n m prev x = function
  | [y] -> median_2 m prev x y
  | _ -> failwith "NonAdditive_c.median_n: binary trees only"
let adjust_3 _ _ t _ _ _ = t, IntSet.empty
let adjust_n _ _ t _ = t, IntSet.empty
let cost t = t.cost
let root_cost t = t.cost
let leaf_cost _ = 0.0
let distance_1 _ a b = float_of_int (distance_ a.eng a.slot b.slot)   (* bv_distance *)
let distance_2 m a b _ = distance_1 m a b
let to_string t =
  let out = Bigarray.Array1.create Bigarray.int8_unsigned Bigarray.c_layout t.n_chars in
  get_states_ t.eng t.slot false out;
  String.concat "," (List.init t.n_chars (fun i -> string_of_int out.{i}))

let of_string _ = failwith "NonAdditive_c.of_string: use create_spec"
let of_parser _ = failwith "NonAdditive_c.of_parser: use create_spec"
let create_spec (e : engine) (chars : codes) n_states (weights : vector option) =
  let n_taxa = Bigarray.Array2.dim1 chars in
  let capacity = 2 * n_taxa in
  set_tips_ e chars n_states weights capacity;
  let s = { engine = e; capacity; next = n_taxa } in
  the_spec := Some s; s
