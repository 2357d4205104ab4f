(* likelihood_c.ml -- body for the reference's empty lib/likelihood_c.ml (all
   `failwith "TODO"`, lib/likelihood_c.ml:1-33), written to the unchanged NodeData.S
   signature (lib/nodeData.ml:3-35) over the B200 engine's C stubs (stubs/phylo_stubs.c).

   NOT COMPILED IN THIS REPOSITORY: the build image has no OCaml toolchain (see DESIGN.md).
   It is kept deliberately mechanical: every function is one external call plus slot
   bookkeeping. A node value [t] names a device-resident CLV by slot; results are fresh
   slots (functional semantics at the OCaml surface, device arena underneath). *)
open Internal

type engine
type matrix = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array2.t
type vector = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type masks  = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t

external engine_create : int -> engine = "phylo_CAML_engine_create"
external set_model_ : engine -> matrix -> matrix -> matrix option
                      -> (vector * vector * vector * float option) -> unit
                    = "likelihood_CAML_set_model"
external set_tips_ : engine -> masks -> vector option -> int -> unit = "likelihood_CAML_set_tips"
external median2_ : engine -> int -> (int * float) -> (int * float) -> unit = "likelihood_CAML_median2"
external edge_lnl_ : engine -> int -> int -> vector -> vector -> unit = "likelihood_CAML_edge_lnl"
external optimize_branch_ : engine -> int -> int -> (float * float * float * float) -> float * float
                           = "likelihood_CAML_optimize_branch"
external edge_eval_ : engine -> int -> int -> vector -> matrix -> unit = "likelihood_CAML_edge_eval"
external score_tree_ :
  engine -> (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array2.t -> matrix
  -> (int * int * float) -> float = "likelihood_CAML_score_tree"

type m = MlModel.t

(* a directed node: its CLV slot and the length of the branch above it *)
type t = {
  eng    : engine;
  slot   : int;
  branch : float;
  codes  : IntSet.t;
  lnl    : float option;   (* set on root nodes *)
}

type spec = { engine : engine; n_taxa : int; capacity : int; mutable next : int; }

let the_spec : spec option ref = ref None
let spec () = match !the_spec with Some s -> s | None -> failwith "Likelihood_c: create_spec first"
let fresh () = let s = spec () in let i = s.next in
  if i >= s.capacity then failwith "Likelihood_c: node capacity exhausted";
  s.next <- i + 1; i

let load_model (e : engine) (model : m) =
  set_model_ e model.MlModel.u model.MlModel.d model.MlModel.ui
    (model.MlModel.priors, model.MlModel.rates, model.MlModel.probs, model.MlModel.pinvar)

(* bookkeeping over character codes *)
let filter_codes set t =
  let c = IntSet.inter set t.codes in
  if IntSet.is_empty c then None else Some { t with codes = c }
let filter_codes_comp set t =
  let c = IntSet.diff t.codes set in
  if IntSet.is_empty c then None else Some { t with codes = c }
let cardinal t = IntSet.cardinal t.codes
let get_codes t = t.codes
let mem codes t = match codes with
  | None -> true
  | Some xs -> List.exists (fun x -> IntSet.mem x t.codes) xs
let union _ a _ = a
let compare a b = Pervasives.compare a.slot b.slot
let recode f t = { t with codes = IntSet.fold (fun x acc -> IntSet.add (f x) acc) t.codes IntSet.empty }

(* kernels *)
let median_1 _ _ x = x
let median_2 _model _prev a b =
  let p = fresh () in
  median2_ a.eng p (a.slot, a.branch) (b.slot, b.branch);
  { a with slot = p; lnl = None; codes = IntSet.union a.codes b.codes }
let median_3 model prev _ a b = median_2 model prev a b
let median_n model prev a = function
  | [b] -> median_2 model prev a b
  | _ -> failwith "Likelihood_c.median_n: binary trees only"

(* readjust node n (= median of a and b) against its third neighbour c: the branch above n
   takes its maximum-likelihood length (device sum table + safeguarded Newton). All of n's
   characters count as changed when the length moved. *)
let adjust_3 _model _codes n _a _b c =
  let t, _lnl = optimize_branch_ n.eng n.slot c.slot (n.branch, 1e-8, 100.0, 1e-8) in
  if abs_float (t -. n.branch) <= 1e-8 *. n.branch then n, IntSet.empty
  else { n with branch = t; lnl = None }, n.codes
let adjust_n model codes n = function
  | [a; b; c] -> adjust_3 model codes n a b c
  | _ -> n, IntSet.empty

(* root edge (a,b) with length t: -lnL (a cost to minimise) *)
let distance_1 _model a b =
  let ts = Bigarray.Array1.of_array Bigarray.float64 Bigarray.c_layout [| a.branch |] in
  let out = Bigarray.Array1.create Bigarray.float64 Bigarray.c_layout 1 in
  edge_lnl_ a.eng a.slot b.slot ts out;
  ~-. (out.{0})
let distance_2 model a b _ = distance_1 model a b
let cost t = match t.lnl with Some l -> ~-. l | None -> 0.0
let root_cost = cost
let leaf_cost _ = 0.0
let to_string t = Printf.sprintf "<clv slot %d, branch %f>" t.slot t.branch

(* not part of NodeData.S *)
let of_string _ = failwith "Likelihood_c.of_string: use of_parser"
let of_parser _ = failwith "Likelihood_c.of_parser: data arrive through create_spec"
let create_spec (model : m) (tips : masks) (weights : vector option) =
  let e = engine_create 0 in
  let n_taxa = Bigarray.Array2.dim1 tips in
  let capacity = 2 * n_taxa in
  load_model e model;
  set_tips_ e tips weights capacity;
  let s = { engine = e; n_taxa; capacity; next = n_taxa } in
  the_spec := Some s; s

(* ---- alignments as text, and all GPUs of the box ------------------------------------------
   [symbol_table] turns an Alphabet.t into the 256-entry table of state sets the engine applies on
   the device (single-character names only; gap and missing fold into "all states" for likelihood,
   lib/mlModel.mli:75-76). [create_spec_text] is create_spec for an alignment that is still the
   characters of the input file. *)
external set_symbol_table_ :
  engine -> (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array1.t option -> unit
  = "phylo_CAML_set_symbol_table"

let symbol_table (alph : Alphabet.t) (n_states : int) =
  let all = Int64.pred (Int64.shift_left 1L n_states) in
  let t = Bigarray.Array1.create Bigarray.int64 Bigarray.c_layout 256 in
  Bigarray.Array1.fill t 0L;
  StringMap.iter
    (fun name code ->
      if String.length name = 1 then begin
        let m = Int64.logand (Int64.of_int code) all in
        let m = if m = 0L then all else m in          (* gap-only / missing: every state *)
        t.{Char.code (Char.uppercase name.[0])} <- m;
        t.{Char.code (Char.lowercase name.[0])} <- m
      end)
    alph.Alphabet.name_code;
  t

let create_spec_text (model : m) (alph : Alphabet.t) (text : masks) (weights : vector option) =
  let e = engine_create 0 in
  let n_taxa = Bigarray.Array2.dim1 text in
  let capacity = 2 * n_taxa in
  load_model e model;
  set_symbol_table_ e (Some (symbol_table alph (Bigarray.Array2.dim1 model.MlModel.u)));
  set_tips_ e text weights capacity;
  let s = { engine = e; n_taxa; capacity; next = n_taxa } in
  the_spec := Some s; s
