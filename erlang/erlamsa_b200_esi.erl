%% erlamsa_b200_esi -- FaaS batch endpoint: one HTTP POST mutates a whole corpus on the GPU (SURVEY.md section 8, row f4).
%%
%% Sits next to the reference's ESI module (src/erlamsa_esi.erl: fuzz/3 and json/3 take ONE sample per request) and is
%% registered the same way (mod_esi, erl_script_alias "/erlamsa" in src/erlamsa_httpsvc.erl): POST /erlamsa/erlamsa_b200_esi/batch
%%
%%   request  (JSON)  {"data": [Base64, ...], "n": N, "skip": K, "seed": "A,B,C", "mutations": "bd,num=3", "patterns": "od,nd",
%%                     "blockscale": 1.0, "gpu": 0}
%%                    data = the corpus; case I (K < I =< N, default N = length(data), K = 0) mutates sample ((I-1) rem length)+1
%%                    with the I-th per-case seed of the parent stream (erlamsa_b200:fuzz_batch/2); seed is mandatory -- an
%%                    unseeded batch would not be reproducible, and the reference's own seed header is unusable anyway
%%                    (erlamsa_esi:parse_headers/2 stores a fun where random:seed/1 wants a tuple)
%%   reply            header erlamsa-status: 0 (+ erlamsa-session), body = JSON array of Base64 strings, one per produced case
%%                    (cases whose result is empty are dropped, as erlamsa_main's record_result/2 does)
%%   errors           erlamsa-status 500 / 401, as erlamsa_esi:fuzz/3 answers them
%%
%% Authentication and the token / session headers are the reference's (erlamsa_cmanager:get_client_context/2, :110-121 there).
%% Option strings go through the reference's own parsers, so "-m" / "-p" mean here what they mean on the command line.
%% Not compiled in CI (no OTP in the build image); executed under the Erlang evaluator by tests/test_erlang_shim.py.
-module(erlamsa_b200_esi).
-export([batch/3, parse_request/1]).

header(Key, Env, Default) ->
    case lists:keyfind(Key, 1, Env) of {Key, V} -> V; false -> Default end.

%% JSON text -> {Corpus, Opts}; keys are matched as strings, never turned into atoms (same care as erlamsa_esi:parse_json/2)
parse_request(Json) ->
    Map = hd(erlamsa_json:tokens_to_erlang(erlamsa_json:tokenize(iolist_to_binary(Json)))),
    true = is_map(Map),
    Corpus = [base64:decode(D) || D <- maps:get("data", Map)],
    true = Corpus =/= [],
    Opts0 = #{paths => [direct], output => return, seed => erlamsa_cmdparse:parse_seed(maps:get("seed", Map))},
    Opts = maps:fold(fun option/3, Opts0, Map),
    {Corpus, Opts}.

option("n", V, Acc) when is_integer(V), V > 0 -> maps:put(n, V, Acc);
option("skip", V, Acc) when is_integer(V), V >= 0 -> maps:put(skip, V, Acc);
option("gpu", V, Acc) when is_integer(V), V >= 0 -> maps:put(gpu_device, V, Acc);
option("blockscale", V, Acc) when is_number(V) -> maps:put(blockscale, V * 1.0, Acc);
option("mutations", V, Acc) ->
    {ok, M} = erlamsa_cmdparse:string_to_actions(V, "mutations", erlamsa_mutations:default([])),
    maps:put(mutations, M, Acc);
option("patterns", V, Acc) ->
    {ok, P} = erlamsa_cmdparse:string_to_actions(V, "patterns", erlamsa_patterns:default()),
    maps:put(patterns, P, Acc);
option("data", _, Acc) -> Acc;
option("seed", _, Acc) -> Acc;
option(Key, _, _) when Key =:= "n"; Key =:= "skip"; Key =:= "gpu"; Key =:= "blockscale" -> erlang:error(badarg);
option(_Unknown, _, Acc) -> Acc.

reply_headers(Status, nil) -> lists:flatten(io_lib:format("erlamsa-status: ~p\r\n\r\n", [Status]));
reply_headers(Status, Session) -> lists:flatten(io_lib:format("erlamsa-status: ~p\r\nerlamsa-session: ~s\r\n\r\n", [Status, Session])).

to_json([]) -> "[]";
to_json([H | T]) -> ["[\"", base64:encode(H), "\"", [[",\"", base64:encode(O), "\""] || O <- T], "]"].

batch(Sid, Env, In) ->
    try
        Auth = erlamsa_cmanager:get_client_context(header(http_erlamsa_token, Env, nil), header(http_erlamsa_session, Env, nil)),
        Session = case Auth of
                      {ok, {S, _Ctx}} -> S;
                      {error, unauth} -> throw(error_unauth);
                      {error, timeout} -> throw(error_timeout)
                  end,
        {Corpus, Opts} = parse_request(In),
        erlamsa_logger:log(info, "Batch request from IP ~s, session ~p: ~p samples",
                           [header(http_x_real_ip, Env, header(remote_addr, Env, nil)), Sid, length(Corpus)]),
        Outs = erlamsa_b200:fuzz_batch(Corpus, Opts),
        mod_esi:deliver(Sid, [reply_headers(0, Session)]),
        mod_esi:deliver(Sid, [to_json(Outs)])
    catch
        error_unauth ->
            mod_esi:deliver(Sid, reply_headers(401, nil)),
            mod_esi:deliver(Sid, <<>>);
        _Class:_Reason ->
            mod_esi:deliver(Sid, reply_headers(500, nil)),
            mod_esi:deliver(Sid, "Invalid input parameters")
    end.
