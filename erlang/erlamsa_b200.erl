%% erlamsa_b200 -- Erlang side of the B200 batch engine (binding over include/erlamsa_b200.h).
%%
%% The build image has no OTP, so this module is not compiled in CI; the NIF underneath is (against a mock erl_nif.h,
%% tests/test_nif_harness.py). See INTEGRATION.md.
%%
%%   erlamsa_b200:fuzz_batch(Corpus :: [binary()], Opts :: map()) -> [binary()]
%%       same option map as erlamsa_main:fuzzer/1 (seed, mutations, patterns, generators, n, skip, blockscale);
%%       case I (skip < I =< n) mutates lists:nth(((I-1) rem length(Corpus)) + 1, Corpus) with the I-th per-case seed of the
%%       parent stream -- for a one-element corpus exactly erlamsa_main:fuzzer(#{paths => [direct], input => B, ...}).
%%   erlamsa_b200:fuzzer(Opts) / fuzz(Opts)
%%       drop-in for erlamsa_main:fuzzer/1: routes paths == [direct] with output == return to the GPU and everything else
%%       (stdin, files, network outputs, external modules, unseeded runs) to the untouched Erlang path.
%%   Option `gpu_device` (default 0) picks the GPU; every device has its own engine context inside the NIF.
-module(erlamsa_b200).
-export([fuzz_batch/2, fuzzer/1, fuzz/1, supported/1]).
-on_load(init/0).

init() ->
    Path = filename:join(code:priv_dir(erlamsa), "erlamsa_b200_nif"),
    case erlang:load_nif(Path, 0) of
        ok -> ok;
        {error, _} -> ok      %% no GPU / no engine: fuzz_batch_nif/10 stays a stub and every call takes the Erlang path
    end.

fuzz_batch_nif(_Blobs, _N, _Seed, _MutaPri, _PatPri, _First, _BlockScale, _Ssrf, _Gens, _Device) -> {error, nif_not_loaded}.

%% priorities in table order, -1 = not selected (the engine's eb200_opts.muta_pri / pat_pri)
pri_vector(Table, Selected) ->
    M = maps:from_list(Selected),
    [maps:get(Code, M, -1) || Code <- Table].

mutator_table() -> [Name || {_, _, _, Name, _} <- erlamsa_mutations:mutations()].
pattern_table() -> [Name || {_, _, Name, _} <- erlamsa_patterns:patterns()].

%% What make_generator_fun/4 (src/erlamsa_gen.erl:204-237) keeps of the generator list when paths == [direct] and an input is given:
%% direct and random. stdin, file and jump are dropped there whatever their priority (so the DEFAULT list is fine); genfuz is
%% dropped too unless an external generator module is set -- then the Erlang path runs. An unknown name fails in the reference,
%% so it goes there as well.
generator_pris(Gens, Opts) ->
    M = maps:from_list(Gens),
    Unknown = maps:keys(M) -- [random, jump, direct, file, genfuz, stdin],
    External = maps:is_key(genfuz, M) andalso maps:get(external_generator, Opts, nil) =/= nil,
    case {Unknown, External} of
        {[], false} -> {ok, {maps:get(direct, M, -1), maps:get(random, M, -1)}};
        _ -> unsupported
    end.

%% `output => return` must be asked for: the reference's default output is "-" (stdout), src/erlamsa_main.erl:245
supported(Opts) ->
    maps:get(paths, Opts, ["-"]) =:= [direct] andalso maps:get(output, Opts, "-") =:= return
        andalso maps:get(external_mutations, Opts, nil) =:= nil andalso maps:get(external_post, Opts, nil) =:= nil
        andalso maps:get(sequence_muta, Opts, false) =:= false andalso is_tuple(maps:get(seed, Opts, nil))
        andalso generator_pris(maps:get(generators, Opts, erlamsa_gen:default()), Opts) =/= unsupported.

%% one case, by the reference, with the seeds the batch semantics give it: case I draws the I-th gen_predictable_seed() of
%% the parent stream (skip => I - 1 makes the reference burn the first I - 1 without running them). This is O(I) cheap
%% draws per call and cannot be made O(1) through the public API: the scores, the snand mask and the generator choice of a
%% run are themselves drawn from the parent stream right after seeding (src/erlamsa_mutations.erl:1313-1314,1390-1395,
%% src/erlamsa_gen.erl:197-198), so re-seeding "closer to case I" would change them.
reference_case(Corpus, Opts, I) ->
    B = lists:nth(((I - 1) rem length(Corpus)) + 1, Corpus),
    erlamsa_main:fuzzer(maps:merge(Opts, #{paths => [direct], output => return, input => B, n => I, skip => I - 1})).

fuzz_batch(Corpus, Opts) when is_list(Corpus), Corpus =/= [] ->
    Seed = maps:get(seed, Opts),
    N = maps:get(n, Opts, length(Corpus)),
    Skip = maps:get(skip, Opts, 0),
    MutaPri = pri_vector(mutator_table(), maps:get(mutations, Opts, erlamsa_mutations:default([]))),
    PatPri = pri_vector(pattern_table(), maps:get(patterns, Opts, erlamsa_patterns:default())),
    {ok, Gens} = generator_pris(maps:get(generators, Opts, erlamsa_gen:default()), Opts),
    {Host, Port} = erlamsa_mutations:get_ssrf_ep(),
    case fuzz_batch_nif(Corpus, N - Skip, Seed, MutaPri, PatPri, Skip + 1, maps:get(blockscale, Opts, 1.0) * 1.0,
                        {iolist_to_binary(Host), Port}, Gens, maps:get(gpu_device, Opts, 0)) of
        {ok, Outs, _Metas} ->
            %% a flagged case (path without a device implementation, capacity limit) is re-run, alone, by the reference
            lists:append([case O of {flagged, I, _St, _Why} -> reference_case(Corpus, Opts, I); <<>> -> []; _ -> [O] end || O <- Outs]);   %% record_result/2 drops empty results
        {error, _Why} ->
            %% NIF not loaded (no GPU) or the engine refused the batch: the reference path, with the caller's n and skip and
            %% the same case numbering -- never a different number of results than the GPU path would give
            case Corpus of
                [_Single] -> erlamsa_main:fuzzer(maps:merge(Opts, #{paths => [direct], output => return, input => hd(Corpus)}));
                _ -> lists:append([reference_case(Corpus, Opts, I) || I <- lists:seq(Skip + 1, N)])
            end
    end.

fuzzer(Opts) ->
    case supported(Opts) of
        true ->
            Input = maps:get(input, Opts),
            Corpus = case is_list(Input) of true -> Input; false -> [Input] end,
            fuzz_batch(Corpus, maps:merge(#{n => 1}, Opts));       %% the reference's default n is 1 (src/erlamsa_main.erl:130)
        false -> erlamsa_main:fuzzer(Opts)
    end.

fuzz(Opts) -> fuzzer(Opts).
