%% erlamsa_b200 -- Erlang side of the B200 batch engine (binding over include/erlamsa_b200.h).
%%
%% NOT COMPILED IN THIS REPOSITORY'S CI (no OTP in the build image); see INTEGRATION.md.
%%
%%   erlamsa_b200:fuzz_batch(Corpus :: [binary()], Opts :: map()) -> [binary()]
%%       same option map as erlamsa_main:fuzzer/1 (seed, mutations, patterns, n, skip, blockscale);
%%       case I mutates lists:nth(((I-1) rem length(Corpus)) + 1, Corpus).
%%   erlamsa_b200:fuzzer(Opts) / fuzz(Opts)
%%       drop-in for erlamsa_main:fuzzer/1: routes paths == [direct] / output == return to the GPU and
%%       everything else (stdin, files, network outputs, external modules) to the untouched Erlang path.
-module(erlamsa_b200).
-export([fuzz_batch/2, fuzzer/1, fuzz/1, supported/1]).
-on_load(init/0).

init() ->
    Path = filename:join(code:priv_dir(erlamsa), "erlamsa_b200_nif"),
    case erlang:load_nif(Path, 0) of
        ok -> ok;
        {error, _} -> ok      %% no GPU / no engine: fuzz_batch_nif/7 stays a stub and we fall back to Erlang
    end.

fuzz_batch_nif(_Blobs, _N, _Seed, _MutaPri, _PatPri, _First, _BlockScale) -> {error, nif_not_loaded}.

%% priorities in table order, -1 = not selected (the engine's eb200_opts.muta_pri / pat_pri)
pri_vector(Table, Selected) ->
    M = maps:from_list(Selected),
    [maps:get(Code, M, -1) || Code <- Table].

mutator_table() -> [Name || {_, _, _, Name, _} <- erlamsa_mutations:mutations()].
pattern_table() -> [Name || {_, _, Name, _} <- erlamsa_patterns:patterns()].

supported(Opts) ->
    maps:get(paths, Opts, ["-"]) =:= [direct] andalso maps:get(output, Opts, return) =:= return
        andalso maps:get(external_mutations, Opts, nil) =:= nil andalso maps:is_key(seed, Opts).

fuzz_batch(Corpus, Opts) when is_list(Corpus) ->
    Seed = maps:get(seed, Opts),
    N = maps:get(n, Opts, length(Corpus)),
    Skip = maps:get(skip, Opts, 0),
    MutaPri = pri_vector(mutator_table(), maps:get(mutations, Opts, erlamsa_mutations:default([]))),
    PatPri = pri_vector(pattern_table(), maps:get(patterns, Opts, erlamsa_patterns:default())),
    case fuzz_batch_nif(Corpus, N - Skip, Seed, MutaPri, PatPri, Skip + 1, maps:get(blockscale, Opts, 1.0) * 1.0) of
        {ok, Outs} ->
            %% a flagged case (path without a device implementation, capacity limit) is re-run, alone, by the reference:
            %% case I draws the I-th gen_predictable_seed() there too (skip => I - 1), so the list stays what
            %% erlamsa_main:fuzzer/1 would have produced
            Redo = fun(I) ->
                       B = lists:nth(((I - 1) rem length(Corpus)) + 1, Corpus),
                       erlamsa_main:fuzzer(maps:merge(Opts, #{paths => [direct], output => return, input => B, n => I, skip => I - 1}))
                   end,
            lists:append([case O of {flagged, I} -> Redo(I); <<>> -> []; _ -> [O] end || O <- Outs]);   %% record_result/2 drops empty results
        {error, _Why} ->                                          %% unsupported mutator, no GPU, ...: the reference path
            lists:append([erlamsa_main:fuzzer(maps:merge(Opts, #{paths => [direct], output => return, input => B, n => 1}))
                          || B <- Corpus])
    end.

fuzzer(Opts) ->
    case supported(Opts) of
        true ->
            Input = maps:get(input, Opts),
            Corpus = case is_list(Input) of true -> Input; false -> [Input] end,
            fuzz_batch(Corpus, maps:merge(#{n => max(1, length(Corpus))}, Opts));
        false -> erlamsa_main:fuzzer(Opts)
    end.

fuzz(Opts) -> fuzzer(Opts).
